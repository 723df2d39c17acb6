"""Run tools/kernel_times.py once per library variant built by tools/build_variants.py (ESLAM_B200_LIB selects the
library) and print the lines that decide between them."""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ("map.loss_backward_q", "map.q_adam_planes", "map.q_build", "map.importance", "map.sample_rays", "map.iteration",
        "trk.iteration", "trk.render_forward_q", "trk.pose_backward_q", "map.render_forward_q")


def main():
    libs = sorted(glob.glob(os.path.join(ROOT, "build_exp", "variants", "libeslam_b200_sc*.so")))
    if not libs:
        raise SystemExit("no variants: run tools/build_variants.py on the CPU box first")
    for lib in libs:
        env = dict(os.environ, ESLAM_B200_LIB=lib)
        res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "kernel_times.py")], env=env,
                             capture_output=True, text=True)
        print("==", os.path.basename(lib), "rc", res.returncode)
        for line in res.stdout.splitlines():
            if any(line.startswith(k.rstrip()) for k in KEEP):
                print("  ", line)
        if res.returncode:
            print(res.stderr[-2000:])


if __name__ == "__main__":
    main()
