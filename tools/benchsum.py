import json,sys
d=json.load(open(sys.argv[1]))
print({k:d[k] for k in ["value","ms_per_step","gpu_launches","clocks"]}, d["e2e"], "roofline", d["roofline"]["frac"], d["roofline_k1"], "ENC", d["encode"]["value"], d["encode"]["roofline"]["frac"], d["encode"]["e2e"], d["encode"]["e2e_text"]["value"], d.get("cpu_baseline"), d.get("cpu_baseline_incremental"), d["config"]["merge_log_matches_cpu_golden"], d["encode"].get("cpu_baseline",{}).get("full_output"), d["encode"].get("full_output"))
