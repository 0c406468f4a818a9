"""Compiles the C restatement of the oracle (oracle_entry.c) into oracle/_build/liboracle.so."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, '_build')
LIB = os.path.join(OUT, 'liboracle.so')
SRC = os.path.join(HERE, 'oracle_entry.c')


def build(force=False):
    os.makedirs(OUT, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    subprocess.run(['/usr/bin/gcc', '-O3', '-march=x86-64-v3', '-fopenmp', '-fPIC', '-shared', SRC, '-o', LIB], check=True)
    return LIB


if __name__ == '__main__':
    print(build(force=True))
