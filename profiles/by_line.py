"""Aggregate an ncu capture's per-instruction counters by CUDA source line (needs -lineinfo + --import-source on):
    python profiles/by_line.py gpurun_out/prof_x.ncu-rep [top_n]
columns: line, warp-instructions executed, share, avg active threads, stall samples, source text"""
import csv
import io
import subprocess
import sys


def by_line(path):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    out = {}
    hdr = None
    cur_file = None
    for r in rows:
        if len(r) == 2 and r[0] == 'File Path':
            cur_file = r[1]
            continue
        if r and r[0] == 'Line No':
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr) or not cur_file or not cur_file.endswith(('.cu', '.cuh')):
            continue
        d = dict(zip(hdr[2:], r[2:]))            # the first two columns are (line, cuda source); 'Source' repeats for SASS
        try:
            line = int(r[0])
            inst = int(d['Instructions Executed'] or 0)
            thr = int(d['Thread Instructions Executed'] or 0)
            samp = int(d['# Samples'] or 0)
        except (ValueError, KeyError):
            continue
        key = (cur_file.split('/')[-1], line)
        e = out.setdefault(key, [0, 0, 0, r[1]])
        e[0] += inst; e[1] += thr; e[2] += samp
    return out


if __name__ == '__main__':
    res = by_line(sys.argv[1])
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    tot = sum(v[0] for v in res.values()) or 1
    tot_s = sum(v[2] for v in res.values()) or 1
    print(f'total warp-instructions {tot}, stall samples {tot_s}')
    for (f, line), (inst, thr, samp, src) in sorted(res.items(), key=lambda kv: -kv[1][2])[:top]:
        print(f'{f}:{line:4d} inst {inst:11d} ({100 * inst / tot:5.1f}%) thr/inst {thr / max(inst, 1):5.1f}  samples {samp:7d} ({100 * samp / tot_s:5.1f}%)  {src.strip()[:90]}')
