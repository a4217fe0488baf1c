#!/usr/bin/env python
"""drop_nal.py in.hevc out.hevc K — removes the K-th coded picture (VCL NAL units with that picture index) from an Annex-B
stream: the pictures that referenced it make the decoder generate a grey reference (hevc_refs.c:538-606)."""
import sys


def main():
    data = open(sys.argv[1], "rb").read()
    k = int(sys.argv[3])
    # split at start codes (00 00 01, optionally preceded by 00)
    pos, i = [], 0
    while True:
        j = data.find(b"\x00\x00\x01", i)
        if j < 0:
            break
        pos.append(j - 1 if j > 0 and data[j - 1] == 0 else j)
        i = j + 3
    pos.append(len(data))
    out, pic = bytearray(), -1
    for a, b in zip(pos[:-1], pos[1:]):
        nal = data[a:b]
        hdr = nal[4] if nal[2] == 0 else nal[3]
        nal_type = (hdr >> 1) & 0x3f
        if nal_type < 32:                       # VCL
            first = (nal[6] if nal[2] == 0 else nal[5]) & 0x80
            if first:
                pic += 1
            if pic == k:
                continue
        out += nal
    open(sys.argv[2], "wb").write(out)


if __name__ == "__main__":
    main()
