#!/usr/bin/env python
"""profiles/traffic.json from the `ncu --set full` captures of scripts/gpu_profiles_r02.sh:
dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel of each bench workload, per input sample of the
launch that was captured (bench.py multiplies by the samples one launch processes for `roofline.traffic`).

    python scripts/make_traffic.py [directory holding the .ncu-rep files, default gpurun_out] [output, default profiles/traffic.json]"""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out")
dst = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles", "traffic.json")
# workload -> (capture, input samples of the captured launch, algorithmic bytes per sample)
CAPS = {
    "cfg3": ("r02_pfb_fm1", 1 << 26, 12.0),
    "cfg3_p8": ("r02_pfb_cl_p8", 1 << 26, 12.0),
    "cfg3_p16": ("r02_pfb_ws_p16", 1 << 26, 12.0),
    "cfg2": ("r02_pfb_tma_cfg2", 1 << 26, 12.0),
    "cfg5": ("r02_pfb_tma_cfg5", 1 << 25, 12.0),
    "cfg1": ("r02_ddc_lone_cfg1", 1 << 24, 8.0 + 12.0 / 96),
    "ddc64": ("r02_ddc_mma2", 1 << 24, 8.0 + 64 * 12.0 / 640),
    "cfg4_16k": ("r02_fft_frame", 1 << 27, 8.0),
}


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, r = rows[0], rows[1], rows[2]

    def get(name):
        v, u = float(r[hdr.index(name)]), units[hdr.index(name)]
        return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    return r[hdr.index("Kernel Name")], get("dram__bytes_read.sum"), get("dram__bytes_write.sum")


doc = {"_doc": "dram__bytes_read.sum + dram__bytes_write.sum per input sample from one `ncu --set full` capture of the named "
               "kernel (bench.py multiplies by the samples one launch processes); written by scripts/make_traffic.py"}
for wl, (cap, n, alg) in CAPS.items():
    rep = os.path.join(src, cap + ".ncu-rep")
    if not os.path.exists(rep):
        print("missing", rep)
        continue
    k, rd, wr = raw(rep)
    doc[wl] = {"kernel": k.split("(")[0], "dram_bytes_per_sample": round((rd + wr) / n, 3),
               "algorithmic_bytes_per_sample": round(alg, 3),
               "source": "%s.ncu-rep: dram read %.1f MB + write %.1f MB over 2^%d input samples" % (cap, rd / 1e6, wr / 1e6,
                                                                                                   n.bit_length() - 1)}
    print(wl, doc[wl])
json.dump(doc, open(dst, "w"), indent=1)
