#!/usr/bin/env python
"""Markdown table of bench.py JSON lines (profiles/r02_bench_lines.jsonl): one row per line."""
import json
import sys


def main(path):
    rows = [json.loads(l) for l in open(path) if l.strip()]
    ref = {r["metric"]: r for r in rows if r.get("impl") == "reference"}
    print("| workload | value (device stages) | e2e (host buffers in / out) | reference CPU (cores) | e2e / CPU |")
    print("|---|---|---|---|---|")
    for r in rows:
        if r.get("impl") == "reference":
            continue
        cb = r.get("cpu_baseline") or (ref.get(r["metric"]) or {}).get("cpu_baseline")
        cpu = f"{cb['value']:.1f} ({cb['cores']})" if cb else "-"
        ratio = f"{r['e2e']['value'] / cb['value']:.1f}x" if cb else "-"
        wl = r["config"]["workload"].split(":")[0] + (" " + r["config"]["workload"].split("[")[1].rstrip("]") if "[" in r["config"]["workload"] else "")
        print(f"| {wl} | {r['value']:.1f} {r['unit']} ({r['ms_per_step']:.1f} ms/step) | {r['e2e']['value']:.1f} ({r['e2e'].get('ms_per_step', 0):.1f} ms) | {cpu} | {ratio} |")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "profiles/r02_bench_lines.jsonl")
