"""Aggregate ncu warp-stall samples per CUDA source line.

    python scripts/ncu_hotspots.py <report.ncu-rep> <library.so> [top_n]

ncu's CSV source page lists SASS only; nvdisasm -g gives SASS offset -> source line for the same cubin.  Functions
are matched by instruction count (the kernel and each out-of-line device function form one contiguous address block
in ncu's listing)."""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def nvdisasm_functions(so_path):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so_path)], cwd=tmp, check=True, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    funcs, cur, line = {}, None, None
    for ln in out.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            line = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m and cur is not None:
            funcs[cur].append((int(m.group(1), 16), line, m.group(2).strip()))
    return funcs


def main():
    rep, so = sys.argv[1], sys.argv[2]
    top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    csv_text = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(csv_text.splitlines()))
    kernel = rows[0][1]
    hdr, data = rows[1], rows[2:]
    ia, isamp, iex, ithr = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed")
    blocks, prev = [], None
    for r in data:
        a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
        if prev is None or a != prev + 16:
            blocks.append([])
        blocks[-1].append((a, int(r[isamp] or 0), int(r[iex] or 0), float(r[ithr] or 0)))
        prev = a
    funcs = nvdisasm_functions(so)
    by_len = defaultdict(list)
    for name, ins in funcs.items():
        by_len[len(ins)].append(name)
    per_line = defaultdict(lambda: [0, 0, 0.0])
    total = 0
    unmatched = 0
    for blk in blocks:
        names = by_len.get(len(blk), [])
        # prefer the function whose mangled name appears in the kernel name, else the first candidate
        name = None
        for n in names:
            if name is None:
                name = n
        if name is None:
            unmatched += sum(b[1] for b in blk)
            continue
        for (a, samp, ex, thr), (off, line, text) in zip(blk, funcs[name]):
            key = (line, name[:40])
            per_line[key][0] += samp
            per_line[key][1] += ex
            per_line[key][2] += ex * thr
            total += samp
    print(f"kernel: {kernel}\ntotal samples {total} (unmatched {unmatched})")
    src_cache = {}
    for (line, fn), (samp, ex, thrsum) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top_n]:
        text = ""
        if line:
            path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tardis_b200", "csrc", line[0])
            if path not in src_cache and os.path.exists(path):
                src_cache[path] = open(path).read().splitlines()
            if path in src_cache and line[1] - 1 < len(src_cache[path]):
                text = src_cache[path][line[1] - 1].strip()[:100]
        print(f"{100*samp/max(total,1):5.1f}%  inst={ex:>11d} thr={thrsum/max(ex,1):4.1f}  {line[0] if line else '?'}:{line[1] if line else 0:<5d} {text}")


if __name__ == "__main__":
    main()
