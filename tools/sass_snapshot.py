"""Per-kernel SASS fingerprint of libesikf_b200.so (instruction text without addresses / encodings): run before and after a
refactor and compare to prove that a kernel's machine code did not change.
  python tools/sass_snapshot.py before.json ; <edit, rebuild> ; python tools/sass_snapshot.py after.json before.json"""
import subprocess, re, sys, json, hashlib
out = subprocess.run(['cuobjdump','-sass','fast_livo2_b200/libesikf_b200.so'],capture_output=True,text=True).stdout
funcs={}; cur=None
for line in out.splitlines():
    m=re.match(r'\s*Function : (\S+)',line)
    if m: cur=m.group(1); funcs[cur]=[]; continue
    if cur is None: continue
    m=re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+(.*?);',line)
    if m: funcs[cur].append(m.group(1).strip())
h={k:(len(v),hashlib.md5('\n'.join(v).encode()).hexdigest()) for k,v in funcs.items()}
json.dump(h,open(sys.argv[1],'w'),indent=0)
print(len(h),'functions')
if len(sys.argv) > 2:
    ref=json.load(open(sys.argv[2]))
    for k in sorted(set(h)|set(ref)):
        a,b=ref.get(k),h.get(k)
        print(('SAME ' if a is not None and b is not None and list(a)==list(b) else 'DIFF ')+k[:90], a and a[0], b and b[0])
