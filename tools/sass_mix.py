#!/usr/bin/env python
"""Instruction mix of the largest basic blocks of a kernel's SASS (cuobjdump -sass output on stdin
or a file): a quick check, without a GPU, of how many FP64-pipe instructions (and how many DFMAs
with three distinct register sources) one unrolled control interval issues.

  cuobjdump -sass -fun <mangled> lib.so | python tools/sass_mix.py [--top 3]
"""
import re
import sys
from collections import Counter

INS = re.compile(r"^\s*/\*([0-9a-f]{4,})\*/\s+(.*?);")


def blocks(lines):
    targets, ins = set(), []
    for ln in lines:
        m = INS.match(ln)
        if not m:
            continue
        addr, txt = int(m.group(1), 16), m.group(2).strip()
        ins.append((addr, txt))
        for t in re.findall(r"0x([0-9a-f]+)", txt):
            if re.search(r"\b(BRA|BSSY|BRX|JMP|CALL|WARPSYNC)", txt):
                targets.add(int(t, 16))
    out, cur = [], []
    for addr, txt in ins:
        if addr in targets and cur:
            out.append(cur); cur = []
        cur.append((addr, txt))
        if re.search(r"\b(BRA|EXIT|RET|BRX|JMP)\b", txt):
            out.append(cur); cur = []
    if cur:
        out.append(cur)
    return out


def regs(ops):
    r = set()
    for o in ops:
        m = re.match(r"[-|~!]*\|?(R\d+)", o.strip())
        if m and m.group(1) != "RZ":
            r.add(m.group(1))
    return r


def src_slots(body, op):
    """[(register or None, has .reuse flag)] per source operand slot."""
    out = []
    for o in body[len(op):].split(",")[1:]:
        o = o.strip()
        m = re.match(r"[-|~!]*\|?(R\d+)(\.reuse)?", o)
        out.append((m.group(1), bool(m.group(2))) if m and m.group(1) != "RZ" else (None, False))
    return out


def mix(block):
    """Counts per opcode; three = DFMAs with three distinct register sources; hit3 = those of them
    that find at least one source in the operand reuse cache (the register was flagged .reuse in the
    same slot by an earlier instruction and no instruction with a REGISTER in that slot came between)."""
    c = Counter()
    three = 0
    hit3 = 0
    cache = {}
    for _, txt in block:
        body = re.sub(r"^@!?U?P\d+\s+", "", txt)
        op = body.split()[0]
        base = op.split(".")[0]
        c[base] += 1
        slots = src_slots(body, op)
        if base == "DFMA":
            rs = {r for r, _ in slots if r}
            if len(rs) == 3:
                three += 1
                if any(r and cache.get(k) == r for k, (r, _) in enumerate(slots)):
                    hit3 += 1
        for k, (r, fl) in enumerate(slots):
            if r is not None:
                cache[k] = r if fl else None
    return c, three, hit3


def main():
    top = 3
    args = sys.argv[1:]
    if "--top" in args:
        top = int(args[args.index("--top") + 1]); args = [a for k, a in enumerate(args) if k not in (args.index("--top"), args.index("--top") + 1)] if False else [a for a in args if not a.isdigit() and a != "--top"]
    lines = open(args[0]).read().splitlines() if args else sys.stdin.read().splitlines()
    bl = sorted(blocks(lines), key=len, reverse=True)[:top]
    for b in bl:
        c, three, reuse3 = mix(b)
        fp64 = c["DFMA"] + c["DADD"] + c["DMUL"] + c["DSETP"]
        print(f"block @{b[0][0]:#x}: {len(b)} instrs, FP64-pipe {fp64} (DFMA {c['DFMA']} of which 3-reg {three} "
              f"[{reuse3} reuse-cache hits], DADD {c['DADD']}, DMUL {c['DMUL']}, DSETP {c['DSETP']}), F2F {c['F2F']}, "
              f"other {len(b) - fp64 - c['F2F']}")
        rest = Counter({k: v for k, v in c.items() if k not in ("DFMA", "DADD", "DMUL", "DSETP", "F2F")})
        print("   other:", dict(rest.most_common(14)))


if __name__ == "__main__":
    try:
        main()
    except BrokenPipeError:
        pass
