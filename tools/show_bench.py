import json,sys
for l in sys.stdin:
    if not l.startswith('{'): continue
    d=json.loads(l)
    print('value',round(d["value"]/1e6,2),'M col-steps/s  ms/step',round(d["ms_per_step"],2),'e2e',round(d["e2e"]["value"]/1e6,2))
    for k,v in d["roofline"]["kernels"].items(): print('   ',k, round(v["ms_per_launch"],3),'ms', round(v["GBps"],1),'GB/s')
    for k,v in d.get("other_launches_ms",{}).items(): print('      (',k, round(v,3),'ms )')
    if "verify" in d: print('    verify: bit_identical', d["verify"].get("bit_identical"), 'mismatching', d["verify"].get("mismatching_elements"))
