"""The node-local MPI shim (src/mpi_shim.cpp -> lib/libmpi_shim.a) and its launcher (bin/sb_mpirun): host-only, no GPU.
It stands in for the MPI the reference's drivers expect (no MPI exists in this image); a multi-rank world exists only
under sb_mpirun, and a foreign launcher's environment is refused instead of silently running N copies of rank 0."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "lib", "libmpi_shim.a")
RUN = os.path.join(ROOT, "bin", "sb_mpirun")


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    if not (os.path.exists(LIB) and os.path.exists(RUN)):
        pytest.skip("lib/libmpi_shim.a or bin/sb_mpirun not built (`make`)")
    out = tmp_path_factory.mktemp("mpi") / "mpi_shim_check"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "include", "mpi_shim"), os.path.join(ROOT, "tests", "cpp", "mpi_shim_check.cpp"),
                           LIB, "-L/usr/local/cuda/lib64", "-lcudart", "-lrt", "-lpthread", "-o", str(out)])  # fmt: skip
    return str(out)


def clean_env():
    return {k: v for k, v in os.environ.items() if k not in ("WORLD_SIZE", "RANK", "LOCAL_RANK", "OMPI_COMM_WORLD_SIZE", "PMI_SIZE", "SLURM_NTASKS")}


@pytest.mark.parametrize("n", [1, 2, 4, 7])
def test_collectives_and_point_to_point(exe, n):
    out = subprocess.run([RUN, "-n", str(n), exe], env=clean_env(), capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and f"mpi_shim_check OK: {n} rank(s)" in out.stdout, out.stdout + out.stderr


def test_one_rank_without_launcher(exe):
    out = subprocess.run([exe], env=clean_env(), capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and "1 rank(s)" in out.stdout, out.stdout + out.stderr


def test_foreign_launcher_is_refused(exe):
    """linked against the shim but started by another launcher, every rank would believe it is rank 0 of 1 and
    compute the whole domain.  MPI_Init must fail loudly instead."""
    out = subprocess.run([exe], env=dict(clean_env(), WORLD_SIZE="4"), capture_output=True, text=True, timeout=60)
    assert out.returncode != 0 and "node-local MPI shim" in out.stderr


def test_a_failing_rank_takes_the_job_down(exe, tmp_path):
    """A rank that dies must not leave the others spinning in a barrier."""
    src = tmp_path / "die.cpp"
    src.write_text('#include <mpi.h>\n#include <cstdlib>\nint main(int c,char**v){MPI_Init(&c,&v);int r;MPI_Comm_rank(MPI_COMM_WORLD,&r);if(r==1)std::exit(3);MPI_Barrier(MPI_COMM_WORLD);MPI_Finalize();return 0;}\n')
    die = tmp_path / "die"
    subprocess.check_call(["g++", "-O1", "-I", os.path.join(ROOT, "include", "mpi_shim"), str(src), LIB, "-L/usr/local/cuda/lib64", "-lcudart", "-lrt", "-lpthread", "-o", str(die)])
    out = subprocess.run([RUN, "-n", "3", str(die)], env=clean_env(), capture_output=True, text=True, timeout=60)
    assert out.returncode == 3
