"""Sample-parallel sharding of clips over the GPUs of one node -- the only parallelism the reference has
(videocrafter/sample_text2video.py:174-188 + lvdm/utils/dist_utils.py:4-19): every rank owns the full weights, draws its
own noise from `seed + rank`, and ONE all-gather collects the decoded clips.  No collective touches the denoising
loop, so scaling is weak.  Backend-agnostic (NCCL over NVLink on the B200 box, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def world():
    return (dist.get_rank(), dist.get_world_size()) if dist.is_available() and dist.is_initialized() else (0, 1)


def clips_for_rank(n_clips, rank, world_size):
    """Global clip indices this rank renders (round-robin, like ceil(n_samples / world) per rank)."""
    return list(range(rank, n_clips, world_size))


def clip_seed(base_seed, clip_index):
    """Global clip i always uses base_seed + i, independent of the number of ranks (reproducible under resharding)."""
    return base_seed + clip_index


def gather_clips(frames_u8):
    """frames_u8: [F, H, W, 3] uint8 of this rank's clip -> list of world_size tensors (rank order) on every rank."""
    rank, ws = world()
    if ws == 1:
        return [frames_u8]
    out = [torch.empty_like(frames_u8) for _ in range(ws)]
    dist.all_gather(out, frames_u8.contiguous())
    return out
