"""Decodes the scheduling control bits of an nvdisasm -hex listing (sm_7x+ 128-bit encoding): per instruction the stall
count, the scoreboard a variable-latency instruction signals on completion (wr) / on operand read (rd) and the
scoreboards it waits for.   nvdisasm -hex -c X.cubin | python tools/sass_ctrl.py <function substring> [opcode filter]"""
import re
import sys


def main():
    want = sys.argv[1]
    filt = sys.argv[2] if len(sys.argv) > 2 else None
    lines = sys.stdin.read().splitlines()
    active, i = False, 0
    ins = re.compile(r"^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* 0x([0-9a-f]{16}) \*/")
    hi_re = re.compile(r"^\s+/\* 0x([0-9a-f]{16}) \*/")
    while i < len(lines):
        ln = lines[i]
        if ln.startswith("\t.section\t.text."):
            active = want in ln
        m = ins.match(ln) if active else None
        if m and i + 1 < len(lines):
            h = hi_re.match(lines[i + 1])
            if h:
                hi = int(h.group(1), 16)
                ctrl = (hi >> 41) & 0x7fffff
                stall, yld = ctrl & 0xf, (ctrl >> 4) & 1
                wr, rd, wait = (ctrl >> 5) & 7, (ctrl >> 8) & 7, (ctrl >> 11) & 0x3f
                txt = m.group(2).strip()
                if filt is None or re.search(filt, txt):
                    print("%s st%-2d %s wr=%s rd=%s wait=%s  %s" % (
                        m.group(1), stall, "Y" if yld else " ", "-" if wr == 7 else wr, "-" if rd == 7 else rd,
                        "".join(str(b) for b in range(6) if wait >> b & 1) or "-", txt))
                i += 1
        elif active and ln.startswith(".L_") and filt is None:
            print(ln)
        i += 1


if __name__ == "__main__":
    main()
