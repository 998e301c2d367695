# developer tool: one-line summary of a bench.py JSON line (sourced by the gpu_run_*.sh scripts)
summ() { python - "$1" "$2" <<'PY'
import sys, json
tag, path = sys.argv[1], sys.argv[2]
try:
    d = json.loads(open(path).read().strip().splitlines()[-1])
    pc = d.get("per_call") or {}
    e = d.get("e2e") or {}
    print(tag, "us/step", round(d["ms_per_step"] * 1e3, 2), "frac", round(d["roofline"]["frac"], 3), "launches", d["gpu_launches"],
          "| per_call", round((pc.get("ms_per_step") or 0) * 1e3, 2), "excl", round((pc.get("ms_per_step_stream_exclusive") or 0) * 1e3, 2), "isolated", pc.get("isolated_launch_us"),
          "| e2e ms", round(e.get("ms_per_step") or 0, 3), "| parity", (d.get("parity") or {}).get("mismatches"), d["config"]["kernel"])
    for k, r in (d.get("sharded") or {}).items():
        print("   sharded", k, "ms/step", round(r["ms_per_step"], 3), "hbm_frac", round(r["hbm_frac"], 3), "parity", r["parity"]["mismatches"], (r.get("compute_roofline") or {}).get("frac"))
except Exception as ex:
    print(tag, "FAILED", ex, open(path).read()[-600:])
PY
}
