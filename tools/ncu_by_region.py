import sys, subprocess, re, collections
rep, ksub = sys.argv[1], sys.argv[2]
out = subprocess.run(["python","tools/ncu_by_line.py",rep,"indelminer_b200/libindelgpu.so",ksub,"2000"],capture_output=True,text=True).stdout.splitlines()
print(out[0])
# region boundaries from source markers
def lines_of(fn):
    return open("indelminer_b200/csrc/"+fn).read().split("\n")
wv = lines_of("warp_vote.cuh")
def find(txt, src=wv):
    for i,l in enumerate(src,1):
        if txt in l: return i
    return 10**9
m1=find("// 1. index the k-mers"); m2=find("// 2. scan the window"); mB=find("const int mine = __popc(take);"); mC=find("// pass C: full chunks"); m3=find("// 3. bin_bands"); m4=find("// 4. leave the table")
lk=find("__device__ __forceinline__ uint32_t kmer_lookup"); pk=find("__device__ __forceinline__ void pack_read_warp"); ka=find("__device__ __forceinline__ uint32_t kmer_at"); ts=find("__device__ __forceinline__ void tab_store"); vh=find("__device__ __forceinline__ void vote_hits_chunk"); sb=find("__device__ __forceinline__ int select_band_warp")
kc = lines_of("kernels.cuh")
d1=find("__device__ void align_diag1", kc); cm=find("__device__ int count_matches", kc); st=find("__device__ int stitch_segments", kc)
inst=collections.Counter(); samp=collections.Counter(); thr=collections.Counter()
for l in out[2:]:
    parts=l.split()
    if len(parts)<5: continue
    f,ln=parts[0].rsplit(":",1); ln=int(ln); i=int(parts[1]); s=int(parts[3])
    if f=="warp_vote.cuh":
        if ln>=m4: r="vote: table cleanup"
        elif ln>=m3: r="vote: select/zero hist"
        elif ln>=mC: r="vote: pass C (dense merge+atomics)"
        elif ln>=mB: r="vote: pass B (append hits)"
        elif ln>=m2: r="vote: pass A (lookups)"
        elif ln>=m1: r="vote: table build"
        elif ln>=sb: r="vote: select_band_warp"
        elif ln>=vh: r="vote: pass C (vote_hits_chunk: merge + atomics)"
        elif ln>=ts: r="vote: tab_store"
        elif ln>=lk: r="vote: kmer_lookup (A+B)"
        elif ln>=ka: r="vote: kmer_at"
        elif ln>=pk: r="pack_read"
        else: r="warp_vote: layout/mbar/bind"
    elif f=="kernels.cuh":
        if ln>=st: r="stitch_segments"
        elif ln>=cm: r="junction (count_matches)"
        elif ln>=d1: r="align_diag1"
        else: r="kernels: base_code etc"
    else: r=f
    inst[r]+=i; samp[r]+=s
ti=sum(inst.values()); tsamp=sum(samp.values())
for r,v in inst.most_common():
    print(f"{r:40s} {v:14d} {100*v/ti:6.2f}% inst ({v/1048576:7.1f}/read) {100*samp[r]/tsamp:6.2f}% samples")
