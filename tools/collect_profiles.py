#!/usr/bin/env python
"""Turn what tools/r2_profile_set.sh left in gpurun_out/ into the committed evidence under profiles/:
bench lines, the launch list, the ncu summary of the full capture, the per-preset table and traffic.json.

    python tools/collect_profiles.py r2z
"""
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def main():
    tag = sys.argv[1]
    for src, dst in ((f"{tag}_bench_n1.json",) * 2, (f"{tag}_bench_reference.json",) * 2, (f"{tag}_collection.json",) * 2,
                     (f"{tag}_collection_cat8192.json",) * 2, (f"{tag}_launches.csv", f"{tag}_launch_list_steady_state.csv")):
        shutil.copy(os.path.join(G, src), os.path.join(P, dst))
    summary = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), os.path.join(G, f"{tag}_kernels_full.ncu-rep")],
                             capture_output=True, text=True, check=True).stdout
    with open(os.path.join(P, f"{tag}_kernels_ncu_summary.txt"), "w") as fh:
        fh.write(summary)
    lines = ["# bench.py --steps 100 --warmup 10 --no-cpu-baseline --task <preset>  (B200, 4096 envs, steady state after a 300-step pre-roll; go2_ts: the default run)",
             "# preset          value M/s   ms/step   e2e M/s   dynamics ms   env ms"]
    for t in "go2 go2_ts go2_cat go2_wtw go2_cts go2_ee go2_dreamwaq tron1_pf tron1_pf_ee".split():
        b = json.load(open(os.path.join(G, f"{tag}_bench_n1.json" if t == "go2_ts" else f"{tag}_bench_{t}.json")))
        lines.append(f"{t:16s} {b['value'] / 1e6:9.2f} {b['ms_per_step']:9.4f} {b['e2e']['value'] / 1e6:9.2f} "
                     f"{b['kernels']['dynamics_step_kernel']['avg_ms']:12.4f} {b['kernels']['env_post_step_kernel']['avg_ms']:9.4f}")
    with open(os.path.join(P, f"{tag}_bench_other_presets.txt"), "w") as fh:
        fh.write("\n".join(lines) + "\n")
    vals = {}
    for blk in summary.split("=" * 100):
        m = re.search(r"void (\w+)", blk)
        if not m:
            continue

        def g(metric):
            mm = re.search(metric + r"\s+([\d.]+) (\w+)", blk)
            return float(mm.group(1)) * {"Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Gbyte": 1e9}[mm.group(2)]
        vals.setdefault(m.group(1), []).append(g("dram__bytes_read.sum") + g("dram__bytes_write.sum"))
    tj = {"_comment": f"dram__bytes_read.sum + dram__bytes_write.sum per launch from profiles/{tag}_kernels_ncu_summary.txt (ncu --set full, go2_ts, 4096 "
                      "envs, steady state; mean of the captured launches of each kernel); read by bench.py for roofline.traffic.  Below the "
                      "algorithmic bytes because the 126 MB L2 absorbs the writes (frames, obs) within the launch",
          "workload": "go2_ts/4096",
          "env_post_step_kernel": int(sum(vals["env_post_step_kernel_preset"]) / len(vals["env_post_step_kernel_preset"])),
          "dynamics_step_kernel": int(sum(vals["dynamics_step_kernel"]) / len(vals["dynamics_step_kernel"]))}
    with open(os.path.join(P, "traffic.json"), "w") as fh:
        json.dump(tj, fh, indent=1)
    print("\n".join(lines))
    b = json.load(open(os.path.join(G, f"{tag}_bench_n1.json")))
    r = json.load(open(os.path.join(G, f"{tag}_bench_reference.json")))
    print("default line:", round(b["value"] / 1e6, 2), b["ms_per_step"], "e2e", round(b["e2e"]["value"] / 1e6, 2), b["e2e"]["ms_per_step"],
          "| reference arm", round(r["value"] / 1e6, 3), "| same config", r["config"] == b["config"])
    print(json.dumps(tj)[-160:])


if __name__ == "__main__":
    main()
