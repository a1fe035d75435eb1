"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import csv, sys, re, collections
def main(path):
    rows=[]
    with open(path) as f:
        lines=[l for l in f if not l.startswith('==')]
    rd=csv.DictReader(lines)
    agg=collections.OrderedDict()
    for r in rd:
        if r.get('Metric Name')!='gpu__time_duration.sum': continue
        name=r['Kernel Name']
        name=re.sub(r'\(.*','',name)
        name=name.replace('lcgp::','').replace('void ','')
        v=float(r['Metric Value'].replace(',',''))
        unit=r['Metric Unit']
        us={'ns':1e-3,'us':1.0,'ms':1e3,'usecond':1.0,'nsecond':1e-3,'msecond':1e3,'second':1e6}.get(unit,1.0)*v
        a=agg.setdefault(name,[0,0.0,0.0]); a[0]+=1; a[1]+=us; a[2]=max(a[2],us)
    tot=sum(a[1] for a in agg.values())
    print(f'{"kernel":60s} {"launches":>8s} {"total_ms":>10s} {"share":>7s} {"avg_us":>10s} {"max_us":>10s}')
    for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1]):
        print(f'{k[:60]:60s} {a[0]:8d} {a[1]/1e3:10.3f} {100*a[1]/tot:6.1f}% {a[1]/a[0]:10.1f} {a[2]:10.1f}')
    print(f'{"TOTAL":60s} {sum(a[0] for a in agg.values()):8d} {tot/1e3:10.3f}')
if __name__=='__main__': main(sys.argv[1])
