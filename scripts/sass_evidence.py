#!/usr/bin/env python
"""Static SASS evidence per kernel of libb200chan.so: counts of the Blackwell-specific mnemonics (TMA loads / stores,
mbarrier, cluster / DSMEM, packed f32x2).  usage: sass_evidence.py [lib] > profiles/rNN_sass_evidence.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "radiocapture_rf_b200/libb200chan.so"
KEYS = ("UBLKCP", "UTMALDG", "UTMASTG", "UTMACMDFLUSH", "SYNCS", "STAS", "MAPA", "UCGABAR", "CCTL", "ERRBAR", "FFMA2", "FADD2",
        "FMUL2", "MUFU", "BAR", "ATOMS", "USETMAXREG", "STG.E.ENL2.256", "RED", "UTMAPF", "UTCHMMA", "UTCBAR", "LDTM", "STTM",
        "UTCATOMSWS")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
dem = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.splitlines()
names = dict(zip(re.findall(r"Function : (\S+)", out), dem))
cur, hist = None, collections.OrderedDict()
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = names.get(m.group(1), m.group(1))
        hist[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m:
        op = m.group(1)
        for k in KEYS:
            if op == k or op.startswith(k + ".") or (k == "STG.E.ENL2.256" and op.startswith(k)):
                hist[cur][k] += 1
print("SASS evidence (cuobjdump -sass %s, sm_100a): static instruction counts per kernel" % lib)
print("UBLKCP = cp.async.bulk (1-D TMA), UTMALDG / UTMASTG = cp.async.bulk.tensor load / store, SYNCS = mbarrier ops,")
print("STAS = st.async (DSMEM store + complete_tx), MAPA = mapa (peer CTA address), UCGABAR = barrier.cluster,")
print("FFMA2/FADD2/FMUL2 = packed f32x2 arithmetic, ATOMS = shared-memory atomic, USETMAXREG = setmaxnreg,")
print("UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / tcgen05.st, UTCATOMSWS = tcgen05.alloc\n")
want = sys.argv[2:] or ["pfb_fm1", "pfb_cl", "pfb_fm_ws", "pfb_fm_tma", "fft_frame", "fft_scan", "fft_cols_tma", "fft_rows", "ddc_lone",
                        "ddc_mma", "ddc_tile", "ddc_post", "post_"]
for name, h in hist.items():
    if not any(w in name for w in want) or not h:
        continue
    print(name[:150])
    print("    " + ", ".join("%s %d" % (k, v) for k, v in sorted(h.items())))
