import re,sys
def parse(path):
    lines=open(path).read().split('\n')
    pat=re.compile(r'^\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/')
    pat2=re.compile(r'^\s+/\* (0x[0-9a-f]{16}) \*/')
    i=0; ins=[]
    while i<len(lines):
        m=pat.match(lines[i])
        if m and i+1<len(lines):
            m2=pat2.match(lines[i+1])
            if m2:
                hi=int(m2.group(1),16)
                ins.append((int(m.group(1),16),m.group(2),(hi>>41)&0xf,(hi>>46)&7,(hi>>49)&7,(hi>>52)&0x3f)); i+=2; continue
        i+=1
    return ins
if __name__=="__main__":
    ins=parse(sys.argv[1]); lo=int(sys.argv[2],16); hi=int(sys.argv[3],16)
    body=[x for x in ins if lo<=x[0]<=hi]
    fp=lambda t: any(t.split()[-0 if not t.startswith('@') else 1].startswith(p) for p in ('DADD','DMUL','DFMA'))
    # region report: contiguous runs of non-FP64-dominated code
    win=40
    for s in range(0,len(body),win):
        seg=body[s:s+win]
        nfp=sum(1 for x in seg if fp(x[1]))
        print(hex(seg[0][0]), "n",len(seg),"fp64",nfp,"stall",sum(x[2] for x in seg), "nonfp_stall", sum(x[2] for x in seg if not fp(x[1])))

def findloop(ins):
    best=None
    for a,t,*_ in ins:
        if 'BRA' in t:
            m=re.search(r'0x([0-9a-f]+)',t)
            if m:
                tgt=int(m.group(1),16)
                if tgt<a and (best is None or a-tgt>best[1]-best[0]): best=(tgt,a)
    return best
def summary(path):
    ins=parse(path); lo,hi=findloop(ins)
    body=[x for x in ins if lo<=x[0]<=hi]
    fp=lambda t: any((t.split()[1] if t.startswith('@') else t.split()[0]).startswith(p) for p in ('DADD','DMUL','DFMA'))
    nfp=sum(1 for x in body if fp(x[1]))
    print(path, hex(lo),hex(hi),"instr",len(body),"fp64",nfp,"static",sum(x[2] for x in body),"nonfp static",sum(x[2] for x in body if not fp(x[1])))
