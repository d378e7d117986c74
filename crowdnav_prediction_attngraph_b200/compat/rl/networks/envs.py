"""rl/networks/envs.py -> make_vec_envs of the CUDA engine (same signature, rl/networks/envs.py:97-140).

Differences a caller can observe: none in the contract (obs dict on `device`, reward CPU float32 [N,1], done
np.bool_ [N], infos); the N environments live on the GPU instead of N worker processes.  With the GST wrapper
(`pretext_wrapper=True`) the predictor weights are read from `config.pred.model_dir/checkpoint/epoch_100.pt`, the
file the reference's wrapper loads (rl/vec_env/vec_pretext_normalize.py:62-79)."""
import os

import numpy as np
import torch

from crowdnav_prediction_attngraph_b200 import vec_env as _ve


def _load_gst_params(config):
    path = os.path.join(config.pred.model_dir, "checkpoint", "epoch_100.pt")
    allow = [(np._core.multiarray.scalar, "numpy.core.multiarray.scalar"), np.dtype, np.dtypes.Float64DType,
             np.dtypes.Float32DType, np.dtypes.Int64DType]
    with torch.serialization.safe_globals(allow):
        ck = torch.load(path, map_location="cpu", weights_only=True)
    return ck["model_state_dict"]


def _dist_shard(num_processes):
    """Under torchrun each rank owns `num_processes` environments of a job with world * num_processes."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return None, 0
    import torch.distributed as dist
    rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return num_processes * world, rank * num_processes


def make_vec_envs(env_name, seed, num_processes, gamma, log_dir, device, allow_early_resets, num_frame_stack=None,
                  config=None, ax=None, test_case=-1, wrap_pytorch=True, pretext_wrapper=False):
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("the CUDA engine needs device 'cuda' (drop --no-cuda); there is no CPU fallback")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    nenv_total, rank_offset = _dist_shard(num_processes)
    if nenv_total is not None:
        device = torch.device("cuda", torch.cuda.current_device())
    gst = _load_gst_params(config) if (pretext_wrapper or env_name == "CrowdSimPredRealGST-v0") else None
    envs = _ve.make_vec_envs(env_name, seed, num_processes, gamma, log_dir, device, allow_early_resets,
                             num_frame_stack=num_frame_stack, config=config, ax=ax, test_case=test_case,
                             wrap_pytorch=wrap_pytorch, pretext_wrapper=pretext_wrapper, nenv_total=nenv_total,
                             rank_offset=rank_offset, gst_params=gst)
    if test_case is not None and test_case >= 0 and num_processes == 1:
        base = envs.env if hasattr(envs, "env") else envs
        base.set_state("case_counter", np.array([test_case], np.uint32))
    return envs


class VecNormalize(object):
    """Name kept for `rl/networks/network_utils.py:6` (isinstance test in get_vec_normalize).  The reference wraps only
    1-D Box observation spaces with it (envs.py:118-123); the crowd environments have Dict observations, so no
    engine object is ever an instance."""


VecPyTorch = _ve.CudaCrowdVecEnv            # the engine's vec env already honours the VecPyTorch contract
