"""Turn the files of tools/collect_profiles.sh (gpurun_out/) into the committed evidence under profiles/:
copies the bench lines / launch lists / raw ncu pages, writes the per-launch ncu table (profiles/r02_ncu_summary.md),
the per-source-line instruction attribution of the two heaviest launches, and profiles/traffic.json (read by bench.py).
    python tools/summarize_profiles.py
Evidence tooling (runs on the build box, no GPU)."""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out")
DST = os.path.join(ROOT, "profiles")

COPY = ["r02_bench_line.json", "r02_event_launches_migan512_bs32.json", "r02_bench_migan256_bs32.json",
        "r02_event_launches_migan256_bs32.json", "r02_bench_migan256_bs32_tc_fast.json", "r02_latency_512.json",
        "r02_latency_256.json", "r02_ncu_launch_list_migan512_bs32.csv", "r02_ncu_tc_enc.raw.csv", "r02_ncu_tc_syn.raw.csv",
        "r02_ncu_ew.raw.csv", "r02_bench_comodgan256_bs16.json", "r02_bench_comodgan256_bs16_tc.json",
        "r02_comodgan_gpu_tests.log", "r02_sass_mnemonics.txt", "r02_smi_before.csv", "r02_smi_after.csv"]

NAMES = {
    "r02_ncu_tc_enc": ["encoder.b512.conv1  STEM 4->64->64 (fused fromrgb)", "encoder.b512.conv2  GEMM 64->128 (pre-split A by TMA)",
                       "encoder.b256.conv1  NHWC 128->128", "encoder.b256.conv2  GEMM 128->256", "encoder.b128.conv1  NHWC 256->256 (two N halves)"],
    "r02_ncu_tc_syn": ["synthesis.b64.conv1  GEMM 512->512 @32^2", "synthesis.b64.conv2  GEMM 512->512 @64^2", "synthesis.b128.conv1  NHWC 512->256 raw",
                       "synthesis.b128.conv2  UP 256->256 + torgb", "synthesis.b256.conv1  NHWC 256->128 raw", "synthesis.b256.conv2  UP 128->128 + torgb",
                       "synthesis.b512.conv1  NHWC 128->64 raw", "synthesis.b512.conv2  UP 64->64 + torgb (image only)"],
    "r02_ncu_ew": None,
}
METRICS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "DRAM rd"), ("dram__bytes_write.sum", "DRAM wr"),
           ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
           ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
           ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem LSU wavefronts %"),
           ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
           ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
           ("smsp__inst_executed.sum", "warp instr"), ("launch__registers_per_thread", "regs")]


def table(name):
    path = os.path.join(SRC, name + ".raw.csv")
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    out = ["| launch | " + " | ".join(m[1] for m in METRICS) + " |", "|---|" + "---|" * len(METRICS)]
    recs = []
    for i, r in enumerate(rows[2:]):
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        label = NAMES[name][i] if NAMES[name] and i < len(NAMES[name]) else d["Kernel Name"].split("(")[0].replace("void ", "").replace("migan::", "")
        cells = []
        for key, _ in METRICS:
            v = d.get(key, "")
            try:
                f = float(v.replace(",", ""))
                cells.append(("%.3g %s" % (f, u.get(key, ""))).strip() if key.startswith(("gpu__time", "dram__bytes")) else "%.4g" % f)
            except ValueError:
                cells.append(v)
        out.append("| %s | %s |" % (label, " | ".join(cells)))
        recs.append((label, d, u))
    return "\n".join(out), recs


def lines_only():
    """On the GPU box: per-source-line instruction / stall attribution of the two heaviest launches -> gpurun_out/r02_ncu_lines_*.txt."""
    dis = os.path.join(SRC, "cub_final", "tc.dis")
    os.makedirs(os.path.dirname(dis), exist_ok=True)
    lib = os.path.join(ROOT, "mi-gan_b200", "lib", "libmigan_b200.so")
    subprocess.run("cd %s && cuobjdump -xelf all %s > /dev/null 2>&1 && nvdisasm -g -c sepconv_tc.sm_100a.cubin > tc.dis" % (os.path.dirname(dis), lib), shell=True)
    for name, idx, tag in (("r02_ncu_tc_enc", 0, "enc_b512_conv1_stem"), ("r02_ncu_tc_enc", 2, "enc_b256_conv1_nhwc"),
                           ("r02_ncu_tc_syn", 7, "syn_b512_conv2_up_torgb")):
        src = os.path.join(SRC, name + ".source.csv")
        if os.path.exists(src) and os.path.exists(dis):
            out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), src, dis, str(idx), "40"], capture_output=True, text=True).stdout
            open(os.path.join(SRC, "r02_ncu_lines_%s.txt" % tag), "w").write(out)
    shutil.rmtree(os.path.dirname(dis), ignore_errors=True)


def main():
    if "--lines-only" in sys.argv:
        return lines_only()
    os.makedirs(DST, exist_ok=True)
    for f in COPY:
        p = os.path.join(SRC, f)
        if os.path.exists(p):
            shutil.copy(p, os.path.join(DST, f))
        else:
            print("missing", f)
    parts = ["# ncu `--set full --clock-control none` at the bench configuration (migan-512, 32 images, second forward)\n",
             "Produced by `tools/collect_profiles.sh` on a B200, read with `ncu -i ... --page raw --csv` (raw pages are the `r02_ncu_*.raw.csv` beside this file).\n"]
    traffic = {}
    for name, title in (("r02_ncu_tc_enc", "sepconv_tc_kernel, encoder launches"), ("r02_ncu_tc_syn", "sepconv_tc_kernel, synthesis launches"),
                        ("r02_ncu_ew", "CUDA-core kernels")):
        if not os.path.exists(os.path.join(SRC, name + ".raw.csv")):
            continue
        t, recs = table(name)
        parts += ["\n## %s\n" % title, t, ""]
        if name == "r02_ncu_tc_enc":
            label, d, u = recs[0]

            def gb(key):
                v = float(d[key].replace(",", ""))
                unit = u[key].lower()
                return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0}.get(unit, 1.0)
            rd, wr = gb("dram__bytes_read.sum"), gb("dram__bytes_write.sum")
            traffic["sepconv_tc"] = {"launch": label + ", 32 images, ncu --set full, profiles/r02_ncu_tc_enc.raw.csv",
                                     "dram_bytes": rd + wr, "dram_read_bytes": rd, "dram_write_bytes": wr,
                                     "alg_bytes": 32.0 * 512 * 512 * (4 + 64) * 4, "duration_under_ncu": d["gpu__time_duration.sum"] + " " + u["gpu__time_duration.sum"]}
    open(os.path.join(DST, "r02_ncu_summary.md"), "w").write("\n".join(parts) + "\n")
    if traffic:
        json.dump(traffic, open(os.path.join(DST, "traffic.json"), "w"), indent=1)
    # per-source-line attribution of the two heaviest launches: made on the GPU box (--lines-only, the source pages are too
    # large to travel), copied here
    for f in sorted(os.listdir(SRC)):
        if f.startswith("r02_ncu_lines_"):
            shutil.copy(os.path.join(SRC, f), os.path.join(DST, f))
    print("profiles/ updated")


if __name__ == "__main__":
    main()
