"""Compare the SASS of every kernel in two `cuobjdump -sass` listings, ignoring addresses and
encodings: `python scripts/sass_diff.py old.sass new.sass`.  Used to prove that a refactor of the
shared stencil bodies leaves the machine code of the validated kernels untouched."""
import re, sys
def funcs(path):
    out={}; cur=None
    for line in open(path):
        m=re.search(r"Function : (\S+)", line)
        if m: cur=m.group(1); out[cur]=[]; continue
        if cur and re.search(r"/\*[0-9a-f]{4}\*/", line):
            out[cur].append(re.sub(r"/\*[0-9a-fx]+\*/","",line).strip())
    return out
a=funcs(sys.argv[1]); b=funcs(sys.argv[2])
for k in a:
    print(("SAME " if a[k]==b.get(k) else "DIFF "), k[:60], len(a[k]), len(b.get(k,[])))
for k in b:
    if k not in a: print("NEW  ", k[:60], len(b[k]))
