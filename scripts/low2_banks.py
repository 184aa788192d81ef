"""Design check for tree_low2.cuh: shared-memory wavefronts per 16-byte access of every stage\nunder candidate swizzles (ideal = 4 per warp instruction).  python scripts/low2_banks.py"""
import itertools
M=64
def bitrev(x,bits):
    r=0
    for i in range(bits): r=(r<<1)|((x>>i)&1)
    return r
def cost(addrs):  # addrs: list of 32 element indices (16B units) for one warp instr
    tot=0
    for qw in range(4):
        a=addrs[8*qw:8*qw+8]
        groups={}
        for x in a:
            if x is None: continue
            groups.setdefault(x%8,set()).add(x)
        tot+=max([len(v) for v in groups.values()] or [0])
    return tot  # ideal 4
def report(name,instrs,swz):
    c=[cost([swz(x) if x is not None else None for x in ins]) for ins in instrs]
    print("  %-28s avg wavefronts/instr %.2f (ideal 4) max %d"%(name,sum(c)/len(c),max(c)))
PLANS={16:(None,4),32:(None,8),64:(None,16),128:(4,8),256:(4,16),512:(8,16),1024:(16,16)}
def run(swz,M):
    # front-end stores
    ins=[[t*32+arr*16+pos for t in range(w*32,w*32+32)] for w in range(M//32) for arr in range(2) for pos in range(16)]
    report("front store",ins,swz)
    N=16
    while N<=8*M:
        P=(M*16//N)//2
        # X stage
        ins=[]
        for h in range(2):
            for w in range(M//32):
                for arr in range(4):
                    for j in range(4):
                        row=[]
                        for t in range(w*32,w*32+32):
                            r=t; p=r//(N//8); gp=r%(N//8); g=h*(N//8)+gp
                            row.append(p*4*N+arr*N+4*g+j)
                        ins.append(row)
        report("N=%d X"%N,ins,swz)
        RP,RM=PLANS[N]
        if RP:
            s=4; ins=[]
            nitems=16*M//RP
            for k in range(nitems//M):
                for w in range(M//32):
                    for q in range(RP):
                        row=[]
                        for t in range(w*32,w*32+32):
                            idx=t+k*M
                            o=idx%4; rest=idx//4; g=rest%(N//(RP*4)); wv=rest//(N//(RP*4))
                            p=wv>>1; which=wv&1
                            row.append(p*4*N+N+which*2*N+g*RP*4+o+q*4)
                        ins.append(row)
            report("N=%d P(%d)"%(N,RP),ins,swz)
        s=N//RM; ins=[]
        nitems=16*M//RM
        for k in range(max(1,nitems//M)):
            for w in range(M//32):
                for q in range(RM):
                    row=[]
                    for t in range(w*32,w*32+32):
                        idx=t+k*M
                        if idx>=nitems: row.append(None); continue
                        o=idx%s; wv=idx//s; p=wv>>1; which=wv&1
                        row.append(p*4*N+N+which*2*N+o+q*s)
                    ins.append(row)
        report("N=%d M(%d) s=%d"%(N,RM,s),ins,swz)
        N*=2
print("swz A: i^((i>>3)&7)"); run(lambda i:i^((i>>3)&7),64)
print("swz B: i^((i>>3)&7)^((i>>6)&7)"); run(lambda i:i^((i>>3)&7)^((i>>6)&7),64)
print("swz C: i^(((i>>3)^(i>>5))&7)"); run(lambda i:i^(((i>>3)^(i>>5))&7),64)
