"""Writes the SASS of the hot kernels to profiles/sass/ (cuobjdump on the objects build.py produced).

    python tools/dump_sass.py [round tag, default r01]
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "econpdes.jl_b200", "build")
OUT = os.path.join(ROOT, "profiles", "sass")

# output name -> regex on the mangled function name (every object of the build is searched)
KERNELS = {
    "k_batched_pdesolve_1d_CampbellCochrane_P4": r"k_batched_pdesolve_1dINS_16CampbellCochraneELi4ELb0ELb0E",
    "k_batched_pdesolve_1d_CampbellCochrane_P4_compact": r"k_batched_pdesolve_1dINS_16CampbellCochraneELi4ELb0ELb1E",
    "k_residual_march_BansalYaron": r"k_residual_marchINS_11BansalYaronELi1ELi2E",
    "k_residual_march_Hjb3dSynth": r"k_residual_marchINS_10Hjb3dSynthELi1ELi2E",
    "k_jacobian_march_BansalYaron": r"k_jacobian_marchINS_11BansalYaronELb0ELi0ELi2E",
    "k_jacobian_march_Hjb3dSynth": r"k_jacobian_marchINS_10Hjb3dSynthELb1ELi0ELi2E",
    "k_spmv": r"k_spmvILi3ELb1EfLb1E",
    "k_spmv_block": r"k_spmv_blockILi2ELb1ELi2E",
    "k_line_zebra_chunk_3d_masked": r"k_line_zebra_chunkILi3ELb1ELi4ELb1E",
    "k_line_zebra_chunk_2d": r"k_line_zebra_chunkILi2ELb0ELi4ELb0E",
    "k_rap_1d": r"k_rap_1dILi3ELb1ELi0E",
    "k_prolong_add": r"k_prolong_addILi3E",
    "k_pcr_factor_lines": r"k_pcr_factor_linesILi2E",   # (the 8 x 8 instantiation is 8 MB of text; the cluster kernel 5 MB)
    "k_normal_onenorm": r"k_normal_onenormILi3ELb1E",
    "k_halo_p2p": r"k_halo_p2p",
}


def main(tag="r02"):
    os.makedirs(OUT, exist_ok=True)
    objs = sorted(os.path.join(OBJ, f) for f in os.listdir(OBJ) if f.endswith(".o"))
    listing = {}
    for path in objs:
        out = subprocess.run(["cuobjdump", "-elf", path], capture_output=True, text=True).stdout
        listing[path] = sorted(set(re.findall(r"\.text\.(_Z\w+)", out)))
    for name, pat in KERNELS.items():
        hit = [(path, m) for path, ms in listing.items() for m in ms if re.search(pat, m)]
        if not hit:
            print("no match for", name)
            continue
        path, names = hit[0][0], [hit[0][1]]
        sass = subprocess.run(["cuobjdump", "-sass", "-fun", names[0], path], capture_output=True, text=True).stdout
        with open(os.path.join(OUT, f"{tag}_{name}.sass"), "w") as f:
            f.write(sass)
        n = len(re.findall(r"^\s+/\*[0-9a-f]{4,6}\*/", sass, re.M))
        print(f"{name}: {names[0][:70]} {n} instructions")


if __name__ == "__main__":
    main(*sys.argv[1:])
