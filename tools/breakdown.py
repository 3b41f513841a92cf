import json, sys
bd=json.load(open(sys.argv[1]))
tot=sum(v['total_ms'] for v in bd.values())
print("total event ms", round(tot,2))
for k,v in bd.items():
    ms=v['total_ms']; n=v['launches']; w=v['work']
    rate = w/(ms/1e3)/1e12 if ms else 0
    if ms/tot > 0.004:
        print(f"{k:34s} n={n:4d} total={ms:7.2f}ms {100*ms/tot:5.1f}%  avg={1e3*ms/n:7.1f}us  rate={rate:8.2f} T/s")
