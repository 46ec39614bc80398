import json
for r in json.load(open("gpurun_out/rng_bench.json")):
    print(r["generator"], r["generators"], r["written_to_hbm"], round(r["gdraws_per_s"], 1), round(r["gb_per_s_written"], 1))
