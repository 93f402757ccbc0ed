"""Markdown tables from bench lines:  python scripts/bench_tables.py <bench.json> [<bench_nN.json> ...]"""
import json
import sys


def load(path):
    lines = [json.loads(l) for l in open(path) if l.startswith("{")]
    return lines[-1]


def fmt(x, d=3):
    return "—" if x is None else (f"{x:.{d}e}" if isinstance(x, float) else str(x))


def legs_of(d):
    head = {"workload": d["config"]["workload"], "value": d["value"], "ms_per_step": d["ms_per_step"], "e2e": d["e2e"], "roofline": d["roofline"],
            "cpu_baseline": d.get("cpu_baseline"), "parity": d.get("parity"), "gpu_launches": d.get("gpu_launches"),
            "cross_rank_check": d.get("cross_rank_check")}
    out = [("3 (headline)", head)]
    for k in ("2", "4", "5"):
        if k in (d.get("configs") or {}):
            out.append((k, d["configs"][k]))
    if d.get("strong") and not d["strong"].get("same_as_headline"):
        out.append(("3 strong", d["strong"]))
    return out


def main():
    for path in sys.argv[1:]:
        d = load(path)
        print(f"\n### {path}: N = {d['n_gpus']}, steps {d['steps']}, warm-up {d['warmup']}, SM clock {d['clocks']['sm_mhz']} MHz, reasons {d['clocks']['reasons']}, bench wall {d.get('bench_wall_s', 0):.0f} s\n")
        print("| config | workload | packets/s resident | ms/step | e2e host buffers | e2e device source (with / without per-packet D2H) | kernel | DRAM GB/s (frac of peak) | issue-active / lanes / occupancy | §8(d) bytes ÷ time (× peak) | CPU port packets/s (threads) | spectrum L2 | max rel err J, ν̄, J_blue, Edotlu | cross-rank |")
        print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
        for name, v in legs_of(d):
            r = v.get("roofline") or {}
            e = v.get("e2e") or {}
            ds = e.get("device_source") or {}
            cpu = v.get("cpu_baseline") or {}
            par = v.get("parity") or {}
            mre = par.get("max_rel_err") or {}
            cr = v.get("cross_rank_check") or {}
            print("| " + " | ".join([
                name, v.get("workload", ""), fmt(v.get("value")), f"{v.get('ms_per_step', 0):.1f}", fmt(e.get("value")),
                f"{fmt(ds.get('value'))} / {fmt((ds.get('fused_spectrum_only') or {}).get('value'))}",
                (r.get("kernel") or "").replace("tb::", ""), f"{fmt(r.get('achieved'), 2)} ({fmt(r.get('frac'), 2)})",
                f"{fmt(r.get('issue_active_pct'), 2)} % / {fmt(r.get('lanes_per_instruction'), 2)} / {fmt(r.get('occupancy_pct'), 2)} %",
                f"{fmt(r.get('survey_8d_GBps'), 2)} ({fmt(r.get('survey_8d_frac_of_peak'), 2)})",
                f"{fmt(cpu.get('value'))} ({cpu.get('cores', '—')})", fmt(par.get("spectrum_l2_vs_oracle")),
                ", ".join(fmt(mre.get(k), 1) for k in ("j", "nu_bar", "j_blue", "edotlu")),
                ("—" if not cr else f"J/ν̄ {fmt(cr.get('j_nubar_max_rel_err'), 1)}, Σ {fmt(cr.get('buffer_sum_rel_err'), 1)}")]) + " |")


main()
