#!/usr/bin/env python
"""Static SASS helper: list the loops of one kernel in a cuobjdump -sass dump with their instruction mix,
and print an address range.  Usage: sass_loops.py dump.sass <function-substring> [lo hi]"""
import collections
import re
import sys


def parse(path, fn):
    txt = open(path).read()
    funcs = re.split(r'\n\s*Function : ', txt)
    for f in funcs[1:]:
        name = f.split('\n')[0]
        if fn in name:
            ins = []
            for l in f.split('\n'):
                m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
                if m:
                    ins.append((int(m.group(1), 16), m.group(2).strip()))
            return name, ins
    raise SystemExit("function not found")


def opcode(t):
    parts = t.split()
    op = parts[1] if parts[0].startswith('@') else parts[0]
    return op.split('.')[0]


def main():
    name, ins = parse(sys.argv[1], sys.argv[2])
    print(name, len(ins))
    if len(sys.argv) > 4:
        lo, hi = int(sys.argv[3], 0), int(sys.argv[4], 0)
        for a, t in ins:
            if lo <= a <= hi:
                print(hex(a), t)
        return
    idx = {a: i for i, (a, t) in enumerate(ins)}
    for i, (a, t) in enumerate(ins):
        m = re.search(r'BRA(?:\.\w+)*\s+(?:!?\w+,\s*)?(0x[0-9a-f]+)', t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= a and tgt in idx:
                body = ins[idx[tgt]:i + 1]
                c = collections.Counter(opcode(x[1]) for x in body)
                fp = c['DFMA'] + c['DMUL'] + c['DADD'] + c['DSETP']
                print("loop %s..%s n=%d fp64=%d mufu=%d lds=%d sts=%d  %s" % (hex(tgt), hex(a), len(body), fp, c['MUFU'], c['LDS'], c['STS'],
                                                                           ' '.join('%s:%d' % kv for kv in c.most_common(8))))


if __name__ == "__main__":
    main()
