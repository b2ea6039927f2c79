"""Parameter block of the hot path: the module-level constants of the reference generators
(SingleCore:20-49, MultiCore:20-52, GPU:21-56, BeamOpt:24-48) under their reference names."""
from __future__ import annotations

import dataclasses


@dataclasses.dataclass(frozen=True)
class BeamOptParams:
    E: float = 200e9                 # Young's modulus (Pa)
    nu: float = 0.3                  # Poisson ratio
    A: float = 0.01                  # cross-section area; enters only the decoupled axial system
    num_nodes: int = 101
    uniform_udl: float = -1000.0     # applied to every element (BeamOpt: -5000)
    I_0: float = 0.5
    max_e: int = 600                 # BeamOpt: num_epochs = 1000
    lr: float = 0.01
    gamma: float = 0.98
    alpha_moment: float = 1e-2
    alpha_shear: float = 1e-2
    tolerance: float = 5e-3          # GPU / BeamOpt: 1e-2
    patience: int = 5                # SC 5, MC 10 (def default, MultiCore:130), GPU 100, BeamOpt 10
    shear_k: float = 0.03            # A_approx = 0.03 * I ** 0.5          (SingleCore:196)
    bending_eps: float = 1e-6        # 2 * E * I + 1e-6                    (SingleCore:195)
    clamp_min: float = 1e-8          # I_tensor.clamp_(min=1e-8)           (SingleCore:208)
    beta1: float = 0.9               # torch.optim.Adam defaults            (SingleCore:166)
    beta2: float = 0.999
    adam_eps: float = 1e-8
    early_stop: bool = True          # False: exactly max_e epochs (benchmark mode)
    zero_last_node: bool = False     # MultiCore:222-223 writes 0.0 for the last node
    num_cases: int = 1               # load cases sharing one I vector (reference: 1)
    max_forces: int = 4              # M_forces_max (BeamOpt: 5)
    solver: int = 0                  # 0 = three-moment (default), 1 = banded LDL^T (OPS_SOLVER_*)

    @property
    def G(self) -> float:            # shear modulus, SingleCore:22
        return self.E / (2 * (1 + self.nu))

    @property
    def num_elements(self) -> int:
        return self.num_nodes - 1

    def replace(self, **kw) -> "BeamOptParams":
        return dataclasses.replace(self, **kw)

    @staticmethod
    def for_script(which: str) -> "BeamOptParams":
        """Effective constants of each reference script ('SC', 'MC', 'GPU', 'BO')."""
        if which == "SC":
            return BeamOptParams(tolerance=5e-3, patience=5)
        if which == "MC":
            return BeamOptParams(tolerance=5e-3, patience=10, zero_last_node=True)
        if which == "GPU":
            return BeamOptParams(tolerance=1e-2, patience=100)
        if which == "BO":
            return BeamOptParams(tolerance=1e-2, patience=10, max_e=1000, uniform_udl=-5000.0, max_forces=5)
        raise ValueError(f"unknown script {which!r}")
