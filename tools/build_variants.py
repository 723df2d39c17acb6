"""Build the library once per codegen variant (the split-compile thread count selects the
variant) into build_exp/variants/, here on the CPU box; tools/variant_times.py then times them on the GPU:

    python tools/build_variants.py
    gpurun -- 'python tools/variant_times.py > gpurun_out/variant_times.txt'
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from myslam_b200.build import SRC, nvcc_path  # noqa: E402

OUT_DIR = os.path.join(ROOT, "build_exp", "variants")
THREADS = (1, 4, 8)


def main():
    os.makedirs(OUT_DIR, exist_ok=True)
    for n in THREADS:
        out = os.path.join(OUT_DIR, f"libeslam_b200_sc{n}.so")
        cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
               "--split-compile", str(n), "-Xcompiler", "-fPIC", "-shared", "-o", out, SRC]
        subprocess.run(cmd, check=True)
        print("built", out)


if __name__ == "__main__":
    main()
