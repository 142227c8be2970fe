import json, sys
for path in sys.argv[1:]:
    d = json.loads(open(path).read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    e = d.get("e2e") or {}
    print(f"{path}: value {d['value']:.0f} q/s  {d['ms_per_step']:.2f} ms/step | e2e {e.get('value', 0):.0f} | gemm {r.get('achieved', 0):.0f} TF/s "
          f"({100*r.get('frac', 0):.1f}% of {r.get('peak')}), gemm {r.get('gemm_ms_per_step', 0):.2f} ms, attn {r.get('attention_ms_per_step', 0):.2f} ms, "
          f"whole-step frac {100*r.get('whole_step_frac', 0):.1f}% | launches {d.get('gpu_launches')} | clocks {d.get('clocks')}")
