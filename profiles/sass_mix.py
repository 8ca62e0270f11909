import subprocess, re, sys, collections
out = subprocess.run(["cuobjdump", "-sass", "gxbeam_b200/libgxbeam_b200.so"], capture_output=True, text=True).stdout
cur=None; counts=collections.defaultdict(collections.Counter); tot=collections.Counter()
for line in out.splitlines():
    m=re.search(r"Function : (\S+)", line)
    if m:
        cur=subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip(); continue
    m=re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op=m.group(1).split('.')[0]; counts[cur][op]+=1; tot[cur]+=1
ops="DMMA DFMA DMUL DADD MUFU UBLKCP UBLKPF SYNCS CREDUX SHFL LDS STS LDG STG LDL STL LDGSTS BAR WARPSYNC".split()
for k in counts:
    if any(s in k for s in sys.argv[1:]):
        print(f"{k}: {tot[k]} SASS instructions; "+", ".join(f"{o} {counts[k][o]}" for o in ops))
