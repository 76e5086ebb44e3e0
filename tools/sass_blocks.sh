#!/bin/bash
# usage: tools/sass_blocks.sh <object> <mangled kernel name> [n]  — prints the n basic blocks with most FFMA2 (run-length op list)
cuobjdump -sass -fun "$2" "$1" | grep -E "^\s+/\*[0-9a-f]{4}\*/" | awk '{ if ($2 ~ /^@/) print $3; else print $2}' | sed 's/\..*//; s/;//' > /tmp/ops.txt
python - "${3:-3}" <<'PY'
import sys
ops=[l.strip() for l in open('/tmp/ops.txt')]
def rl(seq):
    out=[];prev=None;n=0
    for o in seq:
        if o==prev: n+=1
        else:
            if prev: out.append(f"{prev}x{n}" if n>1 else prev)
            prev=o;n=1
    out.append(f"{prev}x{n}")
    return ' '.join(out)
blocks=[];cur=[]
for o in ops:
    cur.append(o)
    if o in ('BRA','EXIT','RET'):
        blocks.append(cur);cur=[]
blocks.append(cur)
print("total instrs", len(ops))
for b in sorted(blocks,key=lambda b:-b.count('FFMA2'))[:int(sys.argv[1])]:
    print(len(b), 'instrs,', b.count('FFMA2'), 'FFMA2'); print(rl(b)); print()
PY
