#!/usr/bin/env python
"""Prints the figures of one bench.py JSON line (last line of the given file) that matter when iterating."""
import json
import sys


def short(v):
    if isinstance(v, float):
        return round(v, 4)
    if isinstance(v, dict):
        return {k: short(x) for k, x in v.items() if k not in ("workload", "note", "kernels_ms_per_step", "hbm_gbs")}
    return v


def main():
    try:
        d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    except Exception as e:          # noqa: BLE001
        print("no bench line in %s: %s" % (sys.argv[1], e))
        return
    r = d.get("roofline") or {}
    print("N=%s %s: %.0f %s, %.4f ms/step; e2e %.0f (%.4f ms)" % (
        d.get("n_gpus"), d["config"]["workload"][:2], d["value"], d["unit"], d["ms_per_step"],
        d["e2e"]["value"], d["e2e"].get("ms_per_step", 0.0)))
    print("  roofline:", r.get("kernel"), short(r.get("frac")), "step", short((r.get("step_contractions") or {}).get("frac")),
          "clocks", d.get("clocks"))
    for k in ("kernels_ms_per_step", "hbm_kernels", "comm", "parity", "peaks"):
        if k in d:
            print("  %s: %s" % (k, json.dumps(short(d[k]) if k != "kernels_ms_per_step" else d[k])))
    for k, v in (d.get("extra") or {}).items():
        print("  extra.%s: %s" % (k, json.dumps(short(v))))


if __name__ == "__main__":
    main()
