"""rayuela_b200 -- B200 (sm_100a) implementation of Rayuela.jl's ICM/ILS encoding and ADC linear scan.

    rayuela_b200.core       memory-image API over librayuela_b200.so (numpy host arrays or torch CUDA tensors)
    rayuela_b200.julia_api  mirror of the Julia package's function names / shapes / index bases
    rayuela_b200.dist       one-process-per-GPU sharding (torch.distributed) of both paths

No CPU fallback exists: importing is cheap, but every call needs the built library and a CUDA device.
"""
from . import core, demos, julia_api, xvecs  # noqa: F401
from .xvecs import bvecs_read, fvecs_read, fvecs_write, ivecs_read, ivecs_write  # noqa: F401
from ._lib import LIB_PATH, RayuelaError, device_count, init, launch_count, shutdown  # noqa: F401
from .julia_api import (get_norms_codebook, quantize_chainq, quantize_norms,  # noqa: F401
                        SR_C_perturb, SR_D_perturb, apply_schedule, encode_icm_cuda, encoding_icm,  # noqa: F401
                        experiment_lsq_cuda, experiment_sr_cuda,
                        eval_recall, linscan_cq, linscan_lsq, linscan_opq, linscan_pq, qerror, qerror_opq,
                        qerror_pq, quantize_opq, quantize_pq, seed_b200, train_lsq, train_lsq_cuda, train_sr_cuda,
                        update_codebooks, update_codebooks_fast_bin, veccost)
