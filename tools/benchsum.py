import json,sys
for f in sys.argv[1:]:
    try:
        d=json.load(open(f)); print(f, "fp/s %.0f ms/step %.2f sad_ms %.2f e2e %.0f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d.get("e2e",{}).get("value",0)))
    except Exception as e: print(f, "ERR", e)
