"""CPU-side tests (no GPU needed): the C-ABI library loads and exports every symbol that
include/b2sv.h declares, error reporting follows the reference's convention, compute entry points
fail loudly without a device (no CPU fallback), and the host-side fusion scheduler produces the
expected plans."""
import ctypes as C
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from cases import layered_circuit, random_circuit, sel_circuit  # noqa: E402
from conftest import HAS_GPU  # noqa: E402


def header_symbols():
    text = open(os.path.join(ROOT, "include", "b2sv.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2sv_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from pennylane_lightning_kokkos_b200 import _lib
    syms = header_symbols()
    assert len(syms) >= 50
    raw = C.CDLL(_lib.LIB_PATH)
    missing = [s for s in syms if not hasattr(raw, s)]
    assert not missing, missing
    # and the ctypes prototypes cover the header
    assert sorted(_lib.PROTOTYPES) == syms


def test_header_cites_the_reference_interface():
    text = open(os.path.join(ROOT, "include", "b2sv.h")).read()
    for cite in ("Bindings.cpp", "SV.hpp", "MK.hpp", "ADJ.hpp", "OBS"):
        assert cite in text


@pytest.mark.parametrize("modname", ["lightning_kokkos_qubit_ops", "lightning_kokkos_qubit_ops_pyb"])
def test_module_surface_matches_reference_binding(modname):
    import importlib
    m = importlib.import_module("pennylane_lightning_kokkos_b200." + modname)
    for bits in ("64", "128"):
        for cls in ("LightningKokkos", "NamedObsKokkos", "HermitianObsKokkos", "TensorProdObsKokkos",
                    "HamiltonianKokkos", "SparseHamiltonianKokkos", "OpsStructKokkos",
                    "AdjointJacobianKokkos"):
            assert hasattr(m, f"{cls}_C{bits}"), cls
        sv = getattr(m, f"LightningKokkos_C{bits}")
        for meth in ("setBasisState", "setStateVector", "apply", "applyGenerator", "ExpectationValue",
                     "probs", "GenerateSamples", "DeviceToHost", "HostToDevice", "numQubits",
                     "dataLength", "resetKokkos", "RX", "CNOT", "DoubleExcitation", "MultiRZ"):
            assert hasattr(sv, meth), meth
    for fn in ("kokkos_start", "kokkos_end", "kokkos_config_info", "print_configuration",
               "InitializationSettings"):
        assert hasattr(m, fn)
    s = m.InitializationSettings().set_device_id(1).set_num_threads(4)
    assert s.get_device_id() == 1 and s.get_num_threads() == 4 and s.has_num_threads()
    # Bindings.cpp:855-866: the constructor calls every setter, so every has_*() is true
    fresh = m.InitializationSettings()
    assert fresh.has_tools_libs() and fresh.has_device_id() and fresh.get_tools_libs() == ""
    assert fresh.get_disable_warnings() is False or fresh.get_disable_warnings() == 0


@pytest.mark.skipif(HAS_GPU, reason="only meaningful on a box without a CUDA device")
def test_no_cpu_fallback():
    from pennylane_lightning_kokkos_b200 import lightning_kokkos_qubit_ops as m
    with pytest.raises(m.PLException) as e:
        m.LightningKokkos_C128(4)
    msg = str(e.value)
    assert "no CUDA device" in msg and "Error in PennyLane Lightning" in msg  # Error.hpp format


def test_ops_struct_errors_match_reference():
    from pennylane_lightning_kokkos_b200 import lightning_kokkos_qubit_ops as m
    with pytest.raises(m.PLException):
        m.OpsStructKokkos_C128(["RX"], [[0.1], [0.2]], [[0]], [False])  # count mismatch
    ops = m.OpsStructKokkos_C128(["RX", "CNOT"], [[0.1], []], [[0], [0, 1]], [False, False])
    assert len(ops) == 2


def plan(circ, n, bits="128"):
    from pennylane_lightning_kokkos_b200 import lightning_kokkos_qubit_ops as m
    cls = getattr(m, f"OpsStructKokkos_C{bits}")
    ops = cls([c[0] for c in circ], [c[3] for c in circ], [c[1] for c in circ], [c[2] for c in circ])
    return ops.plan(n)


def test_scheduler_plan_config2_shape():
    """30-qubit RX/RY/RZ + CNOT-ring layers: 480 gates fuse into ~15 HBM passes, every CNOT is
    folded into the address map, and each wire's RX.RY.RZ triple is pre-multiplied into one 2x2."""
    n, layers = 30, 4
    p = plan(layered_circuit(n, layers, seed=42), n)
    assert p["arithmetic_ops"] == n * layers            # 120 fused 2x2 gates
    # 120 CNOTs, all free, plus the X factors split off 2x2s whose anti-diagonal dominates
    assert n * layers <= p["absorbed_perms"] <= 2 * n * layers
    assert p["passes"] <= 17
    assert p["rounds"] <= 3 * p["passes"]
    assert p["fused_stores"] >= p["passes"] // 2


def test_scheduler_plan_small_and_generic():
    # fewer qubits than a tile: one pass per dependency chain of tile capacity
    p = plan(sel_circuit(8, 2), 8)
    assert p["passes"] >= 1 and p["arithmetic_ops"] >= 16
    # a random circuit over every gate type plans without error for both precisions
    circ = random_circuit(16, 150, seed=9)
    for bits in ("64", "128"):
        q = plan(circ, 16, bits)
        assert q["passes"] >= 1 and q["arithmetic_ops"] > 0
    # identity-only list: nothing to run
    assert plan([("Identity", [0], False, [])], 4)["passes"] == 0


def test_plan_rejects_bad_input():
    from pennylane_lightning_kokkos_b200 import lightning_kokkos_qubit_ops as m
    ops = m.OpsStructKokkos_C128(["RX"], [[0.3]], [[7]], [False])
    with pytest.raises(m.PLException):
        ops.plan(4)  # wire out of range
    ops = m.OpsStructKokkos_C128(["NotAGate"], [[]], [[0]], [False])
    with pytest.raises(m.PLException):
        ops.plan(4)  # neither a named gate nor a matrix


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` is CPU-only: one JSON line with impl/metric/e2e/cpu_baseline; under
    torchrun with 2 ranks only rank 0 prints it and the other rank exits 0."""
    import json
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    bench = os.path.join(root, "bench.py")
    base = [sys.executable, bench, "--impl", "reference", "--ref-qubits", "16", "--steps", "1", "--warmup", "1"]
    out = subprocess.run(base, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    if "unavailable" in d:
        pytest.skip("oracle/_ref not built")
    assert d["impl"] == "reference" and d["unit"] == "GB/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    port = 29700 + (os.getpid() % 1000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(port)] + base[1:] + ["--gpus", "2"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1 and json.loads(lines[0])["impl"] == "reference"
