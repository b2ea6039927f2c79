"""openpystruct_b200 -- B200-native replacement for ONE hot path of dsmyl6/OpenPyStruct:
the per-sample moment-of-inertia optimisation loop of the training-data generators
(OpenPyStruct_BeamOpt_training_{SingleCore,MultiCore,GPU}.py, OpenPyStruct_BeamOpt.py).

Layers (all of them thin; the product is the CUDA kernel behind the C ABI):

  csrc/ + include/openpystruct_b200.h   hand-written sm_100a kernels, ``extern "C"`` boundary
  _cabi                                 ctypes binding of that boundary (no CPU fallback)
  ops                                   ``torch.ops.openpystruct.beam_opt`` custom op (CUDA only)
  sampling / generator                  host mirror of the reference's ``generate_sample`` / ``main``
  distributed                           sample sharding over ranks + one NCCL gather
"""
from .params import BeamOptParams            # noqa: F401
from .generator import (                      # noqa: F401
    GeneratorConfig, generate_sample, generate_samples_batched, generate_dataset, save_training_data,
)

__version__ = "0.1.0"
