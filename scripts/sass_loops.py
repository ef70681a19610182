"""List loops (backward branches) of a cuobjdump -sass dump with their instruction mix."""
import re, sys, collections
lines = open(sys.argv[1]).read().splitlines()
ins = []
for ln in lines:
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr2idx = {a: i for i, (a, _) in enumerate(ins)}
for i, (a, t) in enumerate(ins):
    m = re.search(r'BRA(?:\.\w+)*\s+(?:\w+,\s*)?(0x[0-9a-f]+)', t)
    if m and 'BRA' in t:
        tgt = int(m.group(1), 16)
        if tgt < a and tgt in addr2idx:
            body = ins[addr2idx[tgt]:i + 1]
            if len(body) < 40: continue
            c = collections.Counter()
            for _, x in body:
                op = x.split()[1] if x.startswith('@') else x.split()[0]
                c[op.split('.')[0]] += 1
            fp64 = c['DFMA'] + c['DADD'] + c['DMUL'] + c['DSETP']
            print("loop %05x..%05x n=%d fp64=%d MUFU=%d LDS=%d LDG=%d | %s" % (tgt, a, len(body), fp64, c['MUFU'], c['LDS'], c['LDG'],
                  ' '.join('%s:%d' % kv for kv in c.most_common(14))))
