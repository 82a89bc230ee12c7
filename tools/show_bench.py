"""Print the headline numbers of bench JSON lines: python tools/show_bench.py gpurun_out/r02_bench_n*.json"""
import json, sys
for f in sys.argv[1:]:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    r = d["roofline"]
    print(f"{f}: n_gpus {d['n_gpus']}")
    print(f"  cov {d['value']:.4g} {d['unit']}  {d['ms_per_step']:.2f} ms/step  kernel {r['kernel_ms']:.2f} ms  frac {r['frac']:.3f} (hbm {r['hbm']['frac']:.3f})  e2e {d['e2e']['value']:.3g}")
    for k in ("kriging", "kriging_1deg", "stochastic", "ellipse_fit", "notebook"):
        s = d.get(k)
        if not s:
            continue
        if "error" in s:
            print(f"  {k}: ERROR {s['error']}")
            continue
        e = s.get("e2e") or {}
        f32 = (e.get("float32_error_cov") or {}).get("value")
        print(f"  {k}: {s['value']:.5g} {s['unit']}  ms/step {s.get('ms_per_step', 0):.2f}  frac {(s.get('roofline') or {}).get('frac')}  e2e {e.get('value')}"
              + (f" (float32 E {f32:.0f})" if f32 else "") + f"  cpu {(s.get('cpu_baseline') or {}).get('value')}")
    if d.get("kriging_1deg"):
        print("  golden parity", {k: v for k, v in d["kriging_1deg"].get("golden_parity", {}).items() if k != "what"})
    print("  clocks", d.get("clocks"), " fp64 peaks", {k: round(v, 2) for k, v in (d.get("fp64_peaks") or {}).items() if isinstance(v, float)})
