import re,sys,subprocess
from collections import Counter
obj,pat=sys.argv[1],sys.argv[2]
minop=sys.argv[3] if len(sys.argv)>3 else 'FMUL2'
names=subprocess.run(['cuobjdump','-sass',obj],capture_output=True,text=True).stdout
funs=[l.split('Function : ')[1].strip() for l in names.split('\n') if 'Function :' in l]
fun=[f for f in funs if re.search(pat,f)]
print(fun[:3])
out=subprocess.run(['cuobjdump','-sass','-fun',fun[0],obj],capture_output=True,text=True).stdout
ins=[]
for l in out.split('\n'):
    m=re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);',l)
    if m: ins.append((int(m.group(1),16),m.group(2).strip()))
addr={a:i for i,(a,_) in enumerate(ins)}
print('total',len(ins))
for i,(a,t) in enumerate(ins):
    m=re.search(r'BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)',t)
    if m:
        tgt=int(m.group(1),16)
        if tgt<a and tgt in addr:
            body=ins[addr[tgt]:i+1]
            ops=[re.sub(r'^@!?U?P\d+\s+','',x[1]).split()[0].split('.')[0] for x in body]
            c=Counter(ops)
            if c[minop]>=8: print(hex(tgt),hex(a),len(body),sorted(c.items(),key=lambda x:-x[1]))
