"""Condense the ncu --set full summaries of tools/profile_r02.sh into profiles/<tag>_ncu_traffic.json: measured DRAM bytes
per launch (dram__bytes_read.sum + dram__bytes_write.sum, median over the captured launches) of every kernel bench.py reports
a DRAM-level fraction for, keyed by input set.

    python tools/make_ncu_traffic.py <gpurun_out prefix, e.g. gpurun_out/a2> <profiles tag, e.g. r02>
"""
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PYRAMID_BYTES = 32 * 6 * (64 * 176 + 32 * 88 + 16 * 44 + 8 * 22) * 64 * 4      # zero-fill of the MSMV feature gradients
VALUE_BYTES = 8 * 128 * 128 * 4 * 64 * 4                                          # zero-fill of the MSDA value gradient


def pick(summary, needle):
    for name, e in summary.items():
        if needle in name:
            return e
    return None


def main(prefix, tag):
    out = {"_source": f"profiles/{tag}_ncu_ops_allvalid_summary.json, {tag}_ncu_ops_mixed_summary.json, "
                      f"{tag}_ncu_decoder_summary.json: ncu --set full --clock-control none (tools/profile_r02.sh); ops = the "
                      "tensors of racformer_b200.synthetic.make_op_inputs, decoder = one forward of bench.py's default workload",
           "_note": "backward entries add the size of the zero-fill memset the op performs (a pure DRAM write that is not a "
                    "kernel launch in ncu's list) to the kernel's measured bytes"}
    for case in ("allvalid", "mixed"):
        s = json.load(open(f"{prefix}_ops_{case}_summary.json"))
        shutil.copy(f"{prefix}_ops_{case}_summary.json", os.path.join(ROOT, "profiles", f"{tag}_ncu_ops_{case}_summary.json"))
        out[case] = {"msmv_fwd": pick(s, "msmv_fwd")["dram_traffic_bytes"],
                     "msmv_bwd": pick(s, "msmv_bwd")["dram_traffic_bytes"] + PYRAMID_BYTES,
                     "msda_fwd": pick(s, "msda_fwd")["dram_traffic_bytes"],
                     "msda_bwd": pick(s, "msda_bwd")["dram_traffic_bytes"] + VALUE_BYTES}
    s = json.load(open(f"{prefix}_decoder_summary.json"))
    shutil.copy(f"{prefix}_decoder_summary.json", os.path.join(ROOT, "profiles", f"{tag}_ncu_decoder_summary.json"))
    dec = {}
    for key, needle in (("msmv_fwd", "msmv_fwd"), ("msda_fwd", "msda_fwd"), ("adaptive_mixing_core", "adaptive_mixing_tc"),
                        ("row_programs", "row_program"), ("sasa_attention_core", "sasa_attention")):
        e = pick(s, needle)
        if e and "dram_traffic_bytes" in e:
            dec[key] = e["dram_traffic_bytes"]
    out["decoder_forward_f8"] = dec
    path = os.path.join(ROOT, "profiles", f"{tag}_ncu_traffic.json")
    json.dump(out, open(path, "w"), indent=1)
    print(path)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
