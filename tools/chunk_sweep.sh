# step time vs chunk_rows on ONE box (box-to-box power states differ by ~10 %)
for c in 262144 524288 1048576 262144; do
  NRW_CHUNK_ROWS=$c python bench.py --steps 8 --no_cpu_baseline --no_torch_gpu_ref --no_other_modes 2>/dev/null | tail -1 > /tmp/l.json
  python - "$c" <<'PY'
import sys, json
d = json.load(open("/tmp/l.json"))
print("chunk", sys.argv[1], "ms", round(d["ms_per_step"], 1), "e2e ms", round(d["e2e"]["ms_per_step"], 1), "clk", d["clocks"]["sm_mhz"],
      "slots", d["forward_slots"], "gemm ms", round(d["roofline"]["kernel_ms_per_step"], 1), "launches", d["gpu_launches"])
PY
done
