"""Build libnemoflux_gpu.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m nemoflux_b200.build
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def build(verbose=False, jobs=4):
    cmd = ['make', '-C', os.path.join(HERE, 'csrc'), f'-j{jobs}']
    if not verbose:
        cmd.append('-s')
    subprocess.check_call(cmd)
    return os.path.join(HERE, 'libnemoflux_gpu.so')


if __name__ == '__main__':
    print(build(verbose='-v' in sys.argv))
