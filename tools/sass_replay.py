"""Pipe-level replay of k_line6's SASS for 1-3 warps per scheduler (no GPU needed).

Input: profiles/r2_line6_sass_counts.txt -- one line per SASS instruction of the default shape (build of the round-2 ncu
capture): index, executed warp instructions and stall samples from `ncu --page source` (level 6, 131 072 element pairs),
the control fields decoded from the encoding (st = stall count, y = yield bit, wb / rb = scoreboards set, wm = wait
mask) and the instruction. The replay issues the instructions of one code region (selected by its execution count:
393216 = phase body, 131072 = per-pair part, 262144 = y/x-phase reloads) for N identical warps with round-robin
arbitration, per-pipe issue intervals (FP64 2 cycles, 3 with three register sources; FMA-heavy and ALU pipes 2; MUFU 8;
LDS/STS 2, 4 for 128-bit) and scoreboard latencies. Phase body, two warps: 2303 cycles (measured share of the kernel
time: 2260), FP64 pipe 79 % busy; a third warp would bring 5 %.
    python tools/sass_replay.py [profiles/r2_line6_sass_counts.txt]"""
import re, sys, collections
LAT={'MUFU':18,'LDS':29,'LDC':30,'LDCU':30,'LDG':700,'LDGSTS':30,'S2R':30,'SHFL':25,'DEFAULT':20}
RB_LAT=6
def pipe_of(op, s):
    if op in ('DFMA','DMUL','DADD','DSETP','DMNMX'): return 'fp64'
    if op in ('IMAD','FFMA','FMUL','HFMA2','FADD'): return 'fma'
    if op in ('IADD3','LOP3','SHF','PRMT','ISETP','SEL','FSEL','MOV','VIMNMX','VIMNMX3','LEA','VIADD','FSETP','PLOP3','IABS','FMNMX','CS2R','I2FP','F2FP'): return 'alu'
    if op=='MUFU': return 'xu'
    if op in ('LDS','STS','LDG','STG','LDGSTS','LD','ST','LDL','STL','ATOMS','RED','ATOMG'): return 'lsu'
    if op in ('LDC',): return 'ldc'
    if op.startswith('U') or op in ('LDCU','R2UR'): return 'uni'
    if op in ('BRA','BSSY','BSYNC','CALL','RET','WARPSYNC','EXIT','NOP','BAR','DEPBAR','LDGDEPBAR'): return 'ctl'
    return 'other'
def occupancy(pipe, op, s):
    if pipe=='fp64':
        regs=set(re.findall(r'(?<![U\w])R(\d+)', s.split(',',1)[1] if ',' in s else ''))
        return max(2, len(regs))
    if pipe in ('fma','alu'): return 2
    if pipe=='xu': return 8
    if pipe=='lsu':
        if '.128' in s: return 4
        if '.64' in s: return 2
        return 2
    return 1
R=[]
import os
PATH = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'profiles', 'r2_line6_sass_counts.txt')
for l in open(PATH):
    m=re.match(r'\s*(\d+)\s+(\d+)\s+(\d+) st=\s*(\d+) y=(\d) wb=(\d) rb=(\d) wm=(\w+)\s+(.*)',l)
    R.append(dict(k=int(m.group(1)),ex=int(m.group(2)),sm=int(m.group(3)),st=int(m.group(4)),wb=int(m.group(6)),rb=int(m.group(7)),wm=int(m.group(8),16),s=m.group(9).strip()))
def sim(cnt, nwarps=1):
    sel=[r for r in R if r['ex']==cnt]
    # nwarps identical warps, round-robin arbitration, shared pipes
    W=[dict(pc=0,ready=0,SB=[0]*6) for _ in range(nwarps)]
    pipe_free=collections.defaultdict(int)
    t=0; done=0; busy=collections.Counter()
    off=[i*len(sel)//nwarps for i in range(nwarps)]
    for i,w in enumerate(W): w['pc']=0; w['ready']=0; w['n']=0
    total=len(sel)*nwarps
    last=-1
    while done<total:
        issued=False
        order=list(range(nwarps))
        order=order[(last+1)%nwarps:]+order[:(last+1)%nwarps]
        for wi in order:
            w=W[wi]
            if w['n']>=len(sel): continue
            if w['ready']>t: continue
            r=sel[w['n']]
            op=re.sub(r'^@!?U?P\d+\s+','',r['s']).split()[0].split('.')[0]
            arm=max([w['SB'][i] for i in range(6) if r['wm']>>i &1] or [0])
            if arm>t: continue
            p=pipe_of(op,r['s'])
            if pipe_free[p]>t: continue
            occ=occupancy(p,op,r['s'])
            pipe_free[p]=t+occ; busy[p]+=occ
            if r['wb']<6: w['SB'][r['wb']]=max(w['SB'][r['wb']], t+LAT.get(op,LAT['DEFAULT']))
            if r['rb']<6: w['SB'][r['rb']]=max(w['SB'][r['rb']], t+RB_LAT)
            w['ready']=t+max(r['st'],1); w['n']+=1; done+=1; issued=True; last=wi
            break
        t+=1
    return t, busy
for cnt in (393216,131072,262144):
    for nw in (1,2,3):
        t,b=sim(cnt,nw)
        print(cnt,'warps',nw,'T',t,'per warp-iter',t/nw, {k:round(v/t,2) for k,v in b.items()})
