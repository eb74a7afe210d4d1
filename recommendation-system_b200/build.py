"""Builds libhvae_b200.so (sm_100a) in-tree with nvcc.  `python build.py [--force]`."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OUT = HERE / "hvae_b200" / "libhvae_b200.so"
OBJ = HERE / "build"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "-I", str(HERE.parent / "include"), "-I", str(CSRC), "--expt-relaxed-constexpr"]


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    OBJ.mkdir(exist_ok=True)
    srcs = sorted(CSRC.glob("*.cu"))
    hdrs = list(CSRC.glob("*.cuh")) + [HERE.parent / "include" / "hvae_b200.h"]
    jobs = []
    for s in srcs:
        o = OBJ / (s.stem + ".o")
        if force or _stale(o, [s] + hdrs):
            jobs.append((s, o))

    def cc(job):
        s, o = job
        cmd = [NVCC, *FLAGS, "-c", str(s), "-o", str(o)] + (["-Xptxas", "-v"] if verbose else [])
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {s.name}:\n{r.stdout}\n{r.stderr}")
        return s.name, r.stderr

    with ThreadPoolExecutor(max_workers=8) as ex:
        for name, err in ex.map(cc, jobs):
            if verbose and err:
                print(f"--- {name}\n{err}")
    objs = [OBJ / (s.stem + ".o") for s in srcs]
    if force or jobs or _stale(OUT, objs):
        cmd = [NVCC, "-shared", "-o", str(OUT), *map(str, objs)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
