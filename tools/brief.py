"""Read bench.py JSON lines from stdin and print the few numbers worth eyeballing."""
import json
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else ""
for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    out = [tag, "value=%.4g" % d.get("value", float("nan")), "ms=%.4f" % d.get("ms_per_step", float("nan"))]
    if "roofline" in d:
        out.append("frac=%.3f" % d["roofline"]["frac"])
    if "e2e" in d:
        out.append("e2e=%.4g" % d["e2e"]["value"])
    if "clocks" in d:
        out.append("clk=%s/%s %s" % (d["clocks"].get("sm_mhz"), d["clocks"].get("sm_max_mhz"), d["clocks"].get("reasons")))
    ln = d.get("large_n")
    if ln:
        out.append("large: ms=%.4f frac=%.3f gemv=%.3f upd=%.3f" % (
            ln.get("ms_per_bfgs_step", float("nan")), ln.get("frac_of_peak", float("nan")),
            ln.get("gemv_kernel", {}).get("frac_of_peak", float("nan")),
            ln.get("update_gemv_kernel", {}).get("frac_of_peak", float("nan"))))
    print(" ".join(str(o) for o in out))
