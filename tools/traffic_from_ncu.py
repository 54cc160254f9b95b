"""DRAM traffic per call of the tiled grid kernels from an `ncu --page raw --csv`
export of ONE step (profile_round.sh captures exactly the tiled launches of one
step: collocate launches `tiled_kernel<1,..>`, integrate launches `tiled_kernel<0,..>`).
Usage: python tools/traffic_from_ncu.py raw.csv workload > traffic.json"""
import csv
import json
import sys


def main(path, workload):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    name = col["Kernel Name"]
    rd, wr = col["dram__bytes_read.sum"], col["dram__bytes_write.sum"]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    out = {"collocate": [0.0, 0], "integrate": [0.0, 0]}
    for r in rows[2:]:
        if "tiled_kernel<" not in r[name]:
            continue
        d = "collocate" if "tiled_kernel<1" in r[name] or "tiled_kernel<(bool)1" in r[name] else "integrate"
        out[d][0] += float(r[rd]) * scale[units[rd]] + float(r[wr]) * scale[units[wr]]
        out[d][1] += 1
    print(json.dumps({
        "source": f"{path} (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, summed over the "
                  f"{out['collocate'][1]} collocate / {out['integrate'][1]} integrate tiled launches of one step; 1 GPU)",
        "workload": workload,
        "collocate_bytes_per_call": out["collocate"][0], "integrate_bytes_per_call": out["integrate"][0],
        "collocate_launches": out["collocate"][1], "integrate_launches": out["integrate"][1]}, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "H2O-256")
