"""Back-to-back step time of the default walking-v2 kernel for one library build (ZBOT_B200_LIB = a tuning build of the same
sources, see zbot_lab_b200/build.py):   ZBOT_B200_LIB=... python tools/time_lib.py [envs ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import time_w2  # noqa: E402

if __name__ == "__main__":
    for n in [int(x) for x in sys.argv[1:]] or [65536, 131072]:
        us, name = time_w2.time_one(n, {})
        print(f"{os.path.basename(os.environ.get('ZBOT_B200_LIB', 'libzbot_b200.so')):28s} {n:7d} envs {name:32s} {us:8.2f} us/step "
              f"{n / us * 1e6:.3e} env-steps/s", flush=True)
