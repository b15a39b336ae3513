import collections, csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr=None; cur=None; fname=""
li=collections.Counter(); ls=collections.Counter(); src={}
for r in rows:
    if r and r[0]=="File Path": fname=r[1].split("/")[-1]; continue
    if r and r[0]=="Line No": hdr=r; continue
    if hdr is None or not r or r[0]=="Function Name": continue
    ii, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
    if r[0].strip().isdigit():
        cur=(fname,int(r[0])); src[cur]=r[1].strip()[:110]
        try: li[cur]+=int(r[ii]); ls[cur]+=int(r[isamp])
        except ValueError: pass
tot=sum(li.values()); tots=sum(ls.values())
print("total inst",tot,"samples",tots)
for k,n in ls.most_common(top):
    print(f"{k[0]}:{k[1]:4d} samp {100*n/tots:5.1f}% inst {100*li[k]/tot:5.1f}% | {src[k]}")
