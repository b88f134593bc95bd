"""Per-source-line hot spots of one kernel in an .ncu-rep captured with --import-source on (-lineinfo builds):
warp instructions executed, stall samples and L2 sectors per CUDA source line.

    python profiles/ncu_source_hot.py gpurun_out/x.ncu-rep export_voxel_kernel [top_n]
"""
import csv
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern, "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = None
    lines = []
    cur_file = ""
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) - 5:
            continue
        if r[0] == "":          # SASS line belonging to the previous source line
            continue
        d = dict(zip(hdr, r))
        try:
            lines.append((int(d["Instructions Executed"]), int(d["# Samples"]), cur_file, d["Line No"], d["Source"][1][:110] if False else r[1][:110],
                          d.get("L2 Theoretical Sectors Global", "0"), d.get("stall_long_sb", "0"), d.get("stall_membar", "0")))
        except (ValueError, KeyError):
            continue
    total_inst = sum(l[0] for l in lines)
    total_samp = sum(l[1] for l in lines)
    print(f"kernel {kern}: {total_inst} warp instructions, {total_samp} stall samples")
    print("--- by instructions executed")
    for l in sorted(lines, reverse=True)[:top]:
        print(f"{100.0 * l[0] / max(1, total_inst):5.1f}% inst {100.0 * l[1] / max(1, total_samp):5.1f}% samp  L2sec={l[5]:>10s} longsb={l[6]:>6s} membar={l[7]:>6s} {l[2]}:{l[3]}  {l[4]}")
    print("--- by stall samples")
    for l in sorted(lines, key=lambda x: -x[1])[:top // 2]:
        print(f"{100.0 * l[0] / max(1, total_inst):5.1f}% inst {100.0 * l[1] / max(1, total_samp):5.1f}% samp  L2sec={l[5]:>10s} longsb={l[6]:>6s} membar={l[7]:>6s} {l[2]}:{l[3]}  {l[4]}")


if __name__ == "__main__":
    main()
