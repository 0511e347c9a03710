"""CPU test (-m "not gpu") of the MPI-free start-up behind PetscInitialize (csrc/ohp_boot.cpp): rank / size from the
launcher's environment and the TCP broadcast that carries the 128-byte ncclUniqueId from rank 0 to the other ranks.
No GPU is touched: only ohp_boot_env / ohp_boot_bcast are called."""
import os
import subprocess
import sys
import textwrap

from conftest import ROOT

WORKER = textwrap.dedent("""
    import sys
    sys.path.insert(0, %r)
    import omegahpetsc_b200 as ohp
    rank, n, local = ohp.boot_env()
    payload = bytes((7 * i + 3) %% 256 for i in range(128)) if rank == 0 else bytes(128)
    got = ohp.boot_bcast(rank, n, payload, 128)
    assert got == bytes((7 * i + 3) %% 256 for i in range(128)), (rank, got[:8])
    print("BOOT_OK", rank, n, local)
""")


def test_boot_env_names():
    code = "import sys; sys.path.insert(0, %r); import omegahpetsc_b200 as ohp; print(ohp.boot_env())" % ROOT
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(env, RANK="2", WORLD_SIZE="4", LOCAL_RANK="1"))
    assert out.stdout.strip() == "(2, 4, 1)", out.stdout + out.stderr
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True,
                         env=dict(env, OMPI_COMM_WORLD_RANK="1", OMPI_COMM_WORLD_SIZE="2", OMPI_COMM_WORLD_LOCAL_RANK="1"))
    assert out.stdout.strip() == "(1, 2, 1)", out.stdout + out.stderr
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(env, RANK="5", WORLD_SIZE="4"))
    assert out.returncode != 0 and "rank 5 of 4" in out.stderr


def test_tcp_broadcast_three_ranks(tmp_path):
    script = tmp_path / "boot_worker.py"
    script.write_text(WORKER % ROOT)
    procs = []
    for r in (2, 1, 0):  # rank 0 starts LAST: the others must retry until it listens
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="3", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT="29731",
                   OHP_BOOT_TIMEOUT_S="60")
        procs.append(subprocess.Popen([sys.executable, str(script)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env))
    outs = [p.communicate(timeout=120) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0 and "BOOT_OK" in o, o + e
