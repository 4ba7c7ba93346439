"""Print the main numbers of a bench.py JSON line read from stdin (optional label as argv[1])."""
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
k = d.get("kernels", {})
print(sys.argv[1] if len(sys.argv) > 1 else "", "B/gpu", d["config"].get("angles_per_gpu"), "ms/step %.3f" % d["ms_per_step"],
      "Gpix/s %.2f" % d["value"], "fwd %.3f bwd %.3f" % (k.get("raster_forward_ms", 0), k.get("raster_backward_ms", 0)),
      "e2e", d.get("e2e", {}).get("ms_per_step"), "clk", (d.get("clocks") or {}).get("sm_mhz"), "cand", k.get("candidates"), "batches", k.get("face_batches"))
