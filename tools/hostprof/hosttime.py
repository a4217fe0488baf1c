"""Host-side cost of the drop-in without a GPU: the hooked decoder in record-and-drop mode (B200_SHIM_DUMP=-) against the
unmodified reference decoder, user CPU seconds, best of N runs.  python tools/hostprof/hosttime.py <stream.hevc> [runs] [threads]"""
import os
import resource
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def best(binary, stream, threads, runs, env):
    times = []
    for _ in range(runs):
        before = resource.getrusage(resource.RUSAGE_CHILDREN)
        subprocess.run([os.path.join(ROOT, "oracle", "_ref", binary), stream, threads, "time"], check=True, capture_output=True, env=dict(os.environ, **env))
        after = resource.getrusage(resource.RUSAGE_CHILDREN)
        times.append(after.ru_utime - before.ru_utime + after.ru_stime - before.ru_stime)
    return min(times)


def main():
    stream, runs = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 5
    threads = sys.argv[3] if len(sys.argv) > 3 else "1"
    which = sys.argv[4] if len(sys.argv) > 4 else "both"
    if which != "hooked":
        print(f"reference decoder      : {best('decode_ref', stream, threads, runs, {}):7.3f} CPU s")
    print(f"hooked, record and drop: {best('decode_b200', stream, threads, runs, {'B200_SHIM_DUMP': '-'}):7.3f} CPU s")


main()
