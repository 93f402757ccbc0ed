"""Top stalled SASS instructions of an ncu report, with source line (nvdisasm -g of the matching library).

    python scripts/ncu_sass_hot.py <report.ncu-rep> <library.so> [top_n]
"""
import csv
import subprocess
import sys
from collections import defaultdict

sys.path.insert(0, __import__("os").path.dirname(__file__))
from ncu_hotspots import nvdisasm_functions  # noqa: E402


def main():
    rep, so = sys.argv[1], sys.argv[2]
    top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 60
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, data = rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    blocks, prev = [], None
    for r in data:
        a = int(r[col["Address"]], 16) if r[col["Address"]].startswith("0x") else int(r[col["Address"]])
        if prev is None or a != prev + 16:
            blocks.append([])
        blocks[-1].append(r)
        prev = a
    funcs = nvdisasm_functions(so)
    by_len = defaultdict(list)
    for name, ins in funcs.items():
        by_len[len(ins)].append(name)
    out = []
    total = 0
    for b in blocks:
        names = by_len.get(len(b), [])
        ins = funcs[names[0]] if len(names) >= 1 else None
        for k, r in enumerate(b):
            s = int(r[col["# Samples"]] or 0)
            total += s
            line = ins[k][1] if ins else None
            stalls = sorted(((int(r[col[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
            out.append((s, names[0][:40] if names else "?", k, r[col["Source"]], line, float(r[col["Avg. Threads Executed"]] or 0),
                        int(r[col["Instructions Executed"]] or 0), stalls))
    out.sort(key=lambda x: -x[0])
    print(f"total samples {total}")
    acc = 0
    for s, fn, k, sass, line, thr, nex, stalls in out[:top_n]:
        acc += s
        ln = f"{line[0]}:{line[1]}" if line else "?"
        st = " ".join(f"{n}={v}" for v, n in stalls if v)
        print(f"{100*s/total:5.2f}% cum={100*acc/total:5.1f}% [{fn[-28:]}+{k:5d}] thr={thr:4.1f} n={nex:10d} {ln:28s} {sass[:70]:70s} {st}")


main()
