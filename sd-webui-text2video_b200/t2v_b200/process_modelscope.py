"""`process_modelscope(args_dict, extra_args)` -- the entry point `t2v_helpers.render.run` dispatches to
(reference: scripts/modelscope/process_modelscope.py:34-266), backed by the B200-native pipeline.

Kept: the name, the signature, the module-global `pipe` cache (reset by render.py:41 through `pipe = None`), the
batch loop with `seed + batch` (process_modelscope.py:160,:221) and the argument names of `T2VArgs`
(t2v_helpers/args.py:219-236).  Out of scope by SURVEY.md section 2 rows 5/12: PNG/ffmpeg/base64 packaging, reading /
resizing vid2vid input files, inpainting masks and LoRA UI plumbing (vid2vid itself runs when the frames are passed as a tensor) -- the returned value is therefore the list of clips
(each a list of HxWx3 uint8 BGR frames) rather than data-URL strings, unless a `video_encoder` callable is
installed (e.g. the webui's own ffmpeg_stitch_video wrapper).
"""
from types import SimpleNamespace

import torch

from .pipeline import TextToVideoSynthesis

pipe = None
video_encoder = None        # optional: callable(list_of_bgr_frames, args) -> str (data URL)

_DEFAULTS = dict(prompt='', n_prompt='', steps=30, frames=24, seed=-1, cfg_scale=17, width=256, height=256, eta=0.0,
                 batch_count=1, sampler='DDIM_Gaussian', cpu_vae='GPU (half precision)', keep_pipe_in_vram='None',
                 do_vid2vid=False, model='<modelscope>')          # T2VArgs defaults (t2v_helpers/args.py:219-236)


def process_modelscope(args_dict, extra_args=None):
    """Runs `batch_count` clips.  `args_dict` uses the reference's key names; additionally `model_dir`,
    `prompt_embeds` / `n_prompt_embeds` ([1, L, 1024] tensors) may be given for head-less use."""
    global pipe
    a = SimpleNamespace(**{**_DEFAULTS, **args_dict})
    vid_latents = None
    if getattr(a, 'do_vid2vid', False):
        # The reference reads and resizes the input video with ffmpeg / PIL (process_modelscope.py:118-158): webui plumbing.
        # Head-less use passes the frames as a tensor [1, 3, f, H, W] in [-1, 1]; the latent preparation itself
        # (compute_latents, t2v_pipeline.py:148-194) runs on the library.
        vid = getattr(a, 'vid2vid_frames_tensor', None)
        if vid is None:
            raise NotImplementedError('vid2vid: pass `vid2vid_frames_tensor` ([1, 3, f, H, W] in [-1, 1]); reading / resizing '
                                      'video files is webui plumbing outside this package')
    model_dir = getattr(a, 'model_dir', None)
    if pipe is None or (model_dir is not None and pipe.model_dir != model_dir):
        pipe = TextToVideoSynthesis(model_dir, **(extra_args or {}))
    pipe.keep_in_vram = a.keep_pipe_in_vram
    prompt = getattr(a, 'prompt_embeds', None)
    n_prompt = getattr(a, 'n_prompt_embeds', None)
    prompt = a.prompt if prompt is None else prompt
    n_prompt = a.n_prompt if n_prompt is None else n_prompt
    outputs = []
    if getattr(a, 'do_vid2vid', False):
        vid_latents = pipe.compute_latents(a.vid2vid_frames_tensor, a.cpu_vae, torch.device('cuda')).to(torch.device('cuda'))   # process_modelscope.py:141
    for batch in range(a.batch_count):
        seed = a.seed + batch if a.seed != -1 else -1
        frames, _, info = pipe.infer(prompt, n_prompt, a.steps, a.frames, seed, a.cfg_scale, a.width, a.height, a.eta,
                                     a.cpu_vae, torch.device('cuda'), vid_latents, 0, getattr(a, 'strength', 0.0), None,
                                     vid_latents is not None, a.sampler)
        outputs.append(video_encoder(frames, a) if video_encoder is not None else frames)
    return outputs
