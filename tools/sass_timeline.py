"""Static single-warp timeline of a SASS region (the event-step model of the B300 microarchitecture guide):
T = max(T + stall, scoreboards waited on); variable-latency instructions arm a scoreboard that completes lat cycles later.
Decodes the control bits of sm_70+ 128-bit encodings: stall [105:108], yield 109, wbar [110:112], rbar [113:115],
wait mask [116:121].

    cuobjdump -sass file.o | python tools/sass_timeline.py <function regex> <start addr hex> <end addr hex>
Straight-line estimate: branches are assumed not taken (walks addresses in order), predicated-off instructions still count.
"""
import re
import sys

# variable-latency instructions: measured on B200 with tools/microbench/lat.cu (profiles/r02_microbench_latencies.txt)
LAT = {"LDS": 29, "LDG": 300, "LD": 40, "SHFL": 36, "MUFU": 22, "I2F": 14, "F2I": 14, "I2FP": 14, "POPC": 17, "FLO": 17, "BREV": 17,
       "VOTE": 22, "S2R": 30, "S2UR": 30, "CS2R": 20, "LDC": 40, "ATOMS": 60, "REDUX": 18, "CREDUX": 18, "R2UR": 12, "STS": 6, "ST": 6, "STG": 6,
       "BAR": 30, "LDL": 40, "STL": 6, "MATCH": 34, "DEPBAR": 0}


def parse(stream, fn_re):
    cur, take, out = None, False, []
    pend = None
    for line in stream:
        m = re.search(r"Function : (\S+)", line)
        if m:
            take = re.search(fn_re, m.group(1)) is not None
            continue
        if not take:
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* (0x[0-9a-f]+) \*/", line)
        if m:
            pend = (int(m.group(1), 16), m.group(2).strip(), int(m.group(3), 16))
            continue
        m = re.match(r"\s+/\* (0x[0-9a-f]+) \*/", line)
        if m and pend:
            hi = int(m.group(1), 16)
            out.append((pend[0], pend[1], (hi << 64) | pend[2]))
            pend = None
    return out


def main():
    fn_re, a0, a1 = sys.argv[1], int(sys.argv[2], 16), int(sys.argv[3], 16)
    ins = [i for i in parse(sys.stdin, fn_re) if a0 <= i[0] < a1]
    T = 0
    sb = [0] * 6
    n = 0
    for addr, text, enc in ins:
        stall = (enc >> 105) & 0xF
        wbar = (enc >> 110) & 7
        rbar = (enc >> 113) & 7
        wait = (enc >> 116) & 0x3F
        op = re.sub(r"^@!?U?P\d+\s+", "", text).split()[0].split(".")[0]
        arm = max([sb[k] for k in range(6) if wait >> k & 1] + [0])
        T0 = T
        T = max(T, arm)
        issue = T
        lat = LAT.get(op, 0)
        if wbar < 6:
            sb[wbar] = max(sb[wbar], issue + (lat or 20))
        if rbar < 6:
            sb[rbar] = max(sb[rbar], issue + 6)
        T = issue + max(1, stall)
        n += 1
        print("%05x t=%4d (+%3d wait) stall %2d wb %d rb %d wm %02x  %s" % (addr, issue, issue - T0, stall, wbar, rbar, wait, text))
    print("instructions %d, estimated cycles %d (%.2f per instruction)" % (n, T, T / max(1, n)))


if __name__ == "__main__":
    main()
