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

# output name -> (object, regex on the mangled function name)
KERNELS = {
    "k_batched_pdesolve_1d_CampbellCochrane_P4": ("engine.cu.o", r"k_batched_pdesolve_1dINS_16CampbellCochraneELi4ELb0E"),
    "k_residual_BansalYaron": ("engine.cu.o", r"k_residualINS_11BansalYaron"),
    "k_residual_Hjb3dSynth": ("engine.cu.o", r"k_residualINS_10Hjb3dSynth"),
    "k_jacobian_BansalYaron": ("engine.cu.o", r"k_jacobianINS_11BansalYaronELb0ELi0"),
    "k_jacobian_Hjb3dSynth": ("engine.cu.o", r"k_jacobianINS_10Hjb3dSynthELb1ELi0"),
    "k_spmv": ("krylov.cu.o", r"k_spmvILi3ELb1EfE"),
    "k_spmv_block": ("krylov.cu.o", r"k_spmv_blockILi2ELb1ELi2E"),
    "k_pcr_step_lines": ("krylov.cu.o", r"k_pcr_step_linesILi2E"),
    "k_line_zebra": ("krylov.cu.o", r"k_line_zebraILi3ELb1ELi1E"),
    "k_rap_1d": ("krylov.cu.o", r"k_rap_1dILi3ELb1E"),
    "k_halo_p2p": ("dist.cu.o", r"k_halo_p2p"),
}


def main(tag="r01"):
    os.makedirs(OUT, exist_ok=True)
    for name, (obj, pat) in KERNELS.items():
        path = os.path.join(OBJ, obj)
        funcs = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
        names = [m for m in re.findall(r"Function : (\S+)", funcs) if re.search(pat, m)]
        if not names:
            print("no match for", name)
            continue
        sass = subprocess.run(["cuobjdump", "-sass", "-fun", names[0], path], capture_output=True, text=True).stdout
        with open(os.path.join(OUT, f"{tag}_{name}.sass"), "w") as f:
            f.write(sass)
        n = len(re.findall(r"^\s+/\*[0-9a-f]{4,6}\*/", sass, re.M))
        print(f"{name}: {names[0][:70]} {n} instructions")


if __name__ == "__main__":
    main(*sys.argv[1:])
