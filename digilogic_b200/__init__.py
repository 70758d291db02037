"""digilogic_b200 — B200-native gate-level logic simulation engine behind the gsim builder /
simulator API that rj45/digilogic's `digilogic_gsim` adapter uses.

The product is the C-ABI library (`include/digilogic_b200.h`, sources in `csrc/`): netlist
compiler, HBM-resident bit-sliced wire-state store and hand-written sm_100a kernels.  The
Python modules are the host-side mirror of the reference interface for tests and benchmarks:

* :mod:`digilogic_b200.gsim`    — ``SimulatorBuilder`` / ``Simulator`` / ``LogicState``
* :mod:`digilogic_b200.server`  — ``B200Server`` (the ``SimServer`` implementation, sibling of ``GsimServer``)
* :mod:`digilogic_b200.batch`   — lane-tiled batch runs and stimulus-sharded multi-GPU runs
* :mod:`digilogic_b200.netgen`  — generators for the benchmark netlists
* :mod:`digilogic_b200.importers` — Digital ``.dig`` and Yosys-JSON readers

There is no CPU fallback: without a CUDA device every simulation call raises.
"""
from .gsim import (AddComponentError, B200Error, LogicState, SimulationRunResult, Simulator,  # noqa: F401
                   SimulatorBuilder)

__version__ = "0.1.0"
