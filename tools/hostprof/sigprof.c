/* LD_PRELOAD sampling profiler for the host side of the drop-in (no perf in the image): ITIMER_PROF at 1 kHz, program
 * counters bucketed per mapped object; at exit "object offset count" lines go to $SIGPROF_OUT (default sigprof.txt).
 * tools/hostprof/report.py folds them into functions with nm.  Single- and multi-threaded processes (the timer signal goes to
 * whichever thread is running). */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <signal.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <ucontext.h>

#define MAXS (1 << 22)
static uintptr_t *pcs;
static volatile unsigned long n_pc;

static uintptr_t leaf_lo, leaf_hi;

static void on_prof(int sig, siginfo_t *si, void *uc_)
{
    (void)sig; (void)si;
    ucontext_t *uc = uc_;
    unsigned long i = __sync_fetch_and_add(&n_pc, 1);
    if (i >= MAXS) return;
    uintptr_t pc = (uintptr_t)uc->uc_mcontext.gregs[REG_RIP];
    /* SIGPROF_LEAF=1: a sample inside libc (memset / memcpy: frameless leaves, so [rsp] is the return address) is charged to
     * its caller instead */
    if (leaf_lo && pc >= leaf_lo && pc < leaf_hi) pc = *(const uintptr_t *)uc->uc_mcontext.gregs[REG_RSP];
    pcs[i] = pc;
}

static int cmp(const void *a, const void *b) { uintptr_t x = *(const uintptr_t *)a, y = *(const uintptr_t *)b; return x < y ? -1 : x > y; }

static void dump(void)
{
    struct itimerval off = { { 0, 0 }, { 0, 0 } };
    setitimer(ITIMER_PROF, &off, NULL);
    unsigned long n = n_pc < MAXS ? n_pc : MAXS;
    qsort(pcs, n, sizeof(*pcs), cmp);
    const char *path = getenv("SIGPROF_OUT");
    FILE *f = fopen(path ? path : "sigprof.txt", "w");
    if (!f) return;
    for (unsigned long i = 0; i < n;) {
        unsigned long j = i;
        while (j < n && pcs[j] == pcs[i]) j++;
        Dl_info di;
        if (dladdr((void *)pcs[i], &di) && di.dli_fname)
            fprintf(f, "%s %lx %lu\n", di.dli_fname, (unsigned long)(pcs[i] - (uintptr_t)di.dli_fbase), j - i);
        else
            fprintf(f, "? %lx %lu\n", (unsigned long)pcs[i], j - i);
        i = j;
    }
    fclose(f);
}

__attribute__((constructor)) static void start(void)
{
    pcs = malloc(sizeof(*pcs) * MAXS);
    if (getenv("SIGPROF_LEAF")) {                   /* address range of libc's text, from /proc/self/maps */
        FILE *m = fopen("/proc/self/maps", "r");
        char line[512];
        while (m && fgets(line, sizeof(line), m)) {
            unsigned long lo, hi; char perm[8];
            if (sscanf(line, "%lx-%lx %7s", &lo, &hi, perm) == 3 && strstr(line, "libc.so") && perm[2] == 'x') { leaf_lo = lo; leaf_hi = hi; }
        }
        if (m) fclose(m);
    }
    struct sigaction sa;
    memset(&sa, 0, sizeof(sa));
    sa.sa_sigaction = on_prof;
    sa.sa_flags = SA_SIGINFO | SA_RESTART;
    sigaction(SIGPROF, &sa, NULL);
    struct itimerval it = { { 0, 1000 }, { 0, 1000 } };
    setitimer(ITIMER_PROF, &it, NULL);
    atexit(dump);
}
