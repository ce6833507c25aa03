"""Data side of the pipeline: mirror of the reference's ``VideoDataParser`` (utils/dataparsers/video_dataparser.py:12-156)
and of the frame utilities it uses (utils/VidToMe/utils.py:83-144, 330-346).

Same constructor (``data_config`` with rgb_path / height / width / fps / alpha / flow_model), attributes (``n_frames``,
``unq_inv``, ``fps``, ``alpha``) and methods:

    load_video(frame_ids=None, path=None) -> rgbs [N,3,h,w] in [0,1] on the device
    load_flow(frame_ids, future_flow, past_flow, gts, ...) -> (future_flows, past_flows, mask_bwds)
    load_data(frame_ids=None, rgb_threshold=0.01) -> (rgbs [N*h*w,3], None, None, future_flows, past_flows, mask_bwds)
                                                     and sets ``self.unq_inv``

What differs: frames are decoded with OpenCV (torchvision.io.read_video no longer exists in torchvision 0.26); soft masks,
flow ids and the unique inverse run on the GPU kernels (tclight_b200.flow_utils); and the optical-flow NETWORK
(MemFlow / RAFT) is not part of this package: flows are read from the reference's own cache layout
(``<video>_future_flow_<model>/<frame id:04d>.pt``, written by the reference with ``save_flow=True``) or produced by a
caller-supplied ``flow_fn(src [1,3,H,W], tgt [1,3,H,W]) -> flow [1,2,H,W]``.
"""
from __future__ import annotations

import os
from glob import glob
from typing import Callable, List, Optional

import torch

from ._lib import TclError

FRAME_EXT = [".jpg", ".png", ".jpeg", ".bmp"]


def get_frame_ids(frame_range, num_frames, frame_ids=None):
    """utils/VidToMe/utils.py:330-346."""
    if frame_ids is None:
        frame_range = list(frame_range)
        if len(frame_range) > 1 and frame_range[1] == -1:
            frame_range[1] = num_frames
        if frame_range[1] > num_frames:
            print(f"[WARNING] end frame {frame_range[1]} has been adjusted to number of frames {num_frames}.")
            frame_range[1] = num_frames
        frame_ids = list(range(*frame_range))
    return sorted(frame_ids)


def process_frames(frames: torch.Tensor, h: int, w: int) -> torch.Tensor:
    """Resize so that the frame covers (h, w), then centre-crop (utils/VidToMe/utils.py:83-104,
    utils/general_utils.py:158-179): torchvision's ``Resize`` (bilinear, antialias) + ``CenterCrop`` per frame."""
    import torchvision.transforms as T

    fh, fw = frames.shape[-2:]
    scale = max(w / fw, h / fh)
    size = (int(round(fh * scale)), int(round(fw * scale)))
    if frames.dim() == 3:
        frames = frames[None]
    return torch.stack([T.CenterCrop([h, w])(T.Resize(size)(f)) for f in frames])


def read_frames(path: str) -> torch.Tensor:
    """Video file (.mp4 / .avi / .gif via OpenCV) or a directory of images -> [T,3,H,W] float in [0,1]."""
    import cv2
    import numpy as np

    if os.path.isdir(path):
        paths: List[str] = []
        for ext in FRAME_EXT:
            paths += glob(os.path.join(path, f"*{ext}"))
        paths = sorted(paths)
        if not paths:
            raise TclError(f"no frames found under {path}")
        imgs = [cv2.cvtColor(cv2.imread(p, cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB) for p in paths]
    else:
        cap = cv2.VideoCapture(path)
        if not cap.isOpened():
            raise TclError(f"cannot open video {path}")
        imgs = []
        while True:
            ok, frame = cap.read()
            if not ok:
                break
            imgs.append(cv2.cvtColor(frame, cv2.COLOR_BGR2RGB))
        cap.release()
        if not imgs:
            raise TclError(f"no frames decoded from {path}")
    arr = np.stack(imgs)                                   # [T,H,W,3] uint8
    return torch.from_numpy(arr).permute(0, 3, 1, 2).float() / 255


def load_video(video_path, h, w, frame_ids=None, device="cuda", base=8) -> torch.Tensor:
    """utils/VidToMe/utils.py:115-144."""
    frames = read_frames(video_path)
    if frame_ids is not None:
        frames = frames[list(frame_ids)]
    return process_frames(frames, h, w).to(device)


class VideoDataParser:
    def __init__(self, data_config, device="cuda", dtype=torch.float32, flow_fn: Optional[Callable] = None):
        self.rgb_path = data_config.rgb_path
        self.fps = getattr(data_config, "fps", 30)
        self.alpha = getattr(data_config, "alpha", 0.5)
        self.flow_model = getattr(data_config, "flow_model", "memflow")
        self.h, self.w = data_config.height, data_config.width
        self.voxel_size = None
        self.device = device
        self.dtype = dtype
        self.unq_inv = None
        self.flow_fn = flow_fn
        self.n_frames = self._count_frames()

    def _count_frames(self) -> int:
        if os.path.isdir(self.rgb_path):
            return len([n for n in os.listdir(self.rgb_path) if os.path.isfile(os.path.join(self.rgb_path, n))])
        import cv2

        cap = cv2.VideoCapture(self.rgb_path)
        n = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
        cap.release()
        return n

    @torch.no_grad()
    def load_video(self, frame_ids=None, path=None):
        rgbs = load_video(path or self.rgb_path, self.h, self.w, frame_ids=frame_ids, device=self.device, base=8)
        if rgbs.min() < 0:
            rgbs = (rgbs + 1.0) * 127.0 / 255.0
        self.n_frames = rgbs.shape[0]
        return rgbs

    # ---- flows ------------------------------------------------------------------------------------------
    def create_folder(self, name: str) -> str:
        """video_dataparser.py:126-130: ``<video without ext>_<name>`` or ``<frame dir>/<name>``."""
        ext = os.path.splitext(self.rgb_path)[-1]
        base = self.rgb_path.replace(ext, f"_{name}") if not os.path.isdir(self.rgb_path) else os.path.join(self.rgb_path, name)
        os.makedirs(base, exist_ok=True)
        return base

    def process_flow(self, flow_list):
        """video_dataparser.py:132-137: resize/crop like the frames and scale the vectors by the resize factor."""
        flow = torch.stack(flow_list)
        _, _, H, W = flow.shape
        flow = process_frames(flow, self.h, self.w)
        return flow * max(self.w / W, self.h / H)

    def _one_flow(self, idx, gts, frame_ids, is_future, path, save):
        fname = os.path.join(path, f"{frame_ids[idx]:04d}.pt")
        # video_dataparser.py:115: a cached flow is only trusted when the cache holds exactly this frame selection
        if os.path.exists(fname) and len(os.listdir(path)) == len(frame_ids):
            return torch.load(fname, map_location="cpu")
        zero_idx = gts.shape[0] - 1 if is_future else 0
        src = gts[idx:idx + 1]
        if idx == zero_idx:
            flow = torch.zeros_like(src[:, :2])
        else:
            if self.flow_fn is None:
                raise TclError(f"flow {fname} is not cached and no flow_fn was given: the optical-flow network "
                               "(MemFlow / RAFT) is outside this package (SURVEY.md §8f rank 4)")
            tgt = gts[idx + 1:idx + 2] if is_future else gts[idx - 1:idx]
            flow = self.flow_fn(src, tgt)
        if save:
            torch.save(flow.cpu(), fname)
        return flow

    @torch.no_grad()
    def load_flow(self, frame_ids=None, future_flow=False, past_flow=False, gts=None, target_ids=None, save_flow=True,
                  diff_threshold=0.1):
        """video_dataparser.py:64-110.  ``gts`` are the frames in [0,1]; MemFlow works on 2*gts-1 and the soft masks
        are computed on that range too (:78-79, 107-108)."""
        from . import flow_utils

        frame_ids = list(frame_ids) if frame_ids is not None else list(range(len(gts)))
        if self.flow_model.lower() == "memflow":
            gts = gts * 2.0 - 1.0
        gts = gts.to(self.device, dtype=self.dtype)
        fpath = self.create_folder(f"future_flow_{self.flow_model.lower()}")
        ppath = self.create_folder(f"past_flow_{self.flow_model.lower()}")
        flows, pasts = [], []
        for idx in range(len(gts)):
            if future_flow:
                flows.append(self._one_flow(idx, gts, frame_ids, True, fpath, save_flow)[0].cpu())
            if past_flow:
                pasts.append(self._one_flow(idx, gts, frame_ids, False, ppath, save_flow)[0].cpu())
        flows_t = self.process_flow(flows).to(self.device, dtype=self.dtype) if future_flow else None
        pasts_t = self.process_flow(pasts).to(self.device, dtype=self.dtype) if past_flow else None
        masks = None
        if future_flow and past_flow:
            masks = flow_utils.get_soft_mask_bwds(gts, flows_t, pasts_t, alpha=self.alpha, diff_threshold=diff_threshold)
        return flows_t, pasts_t, masks

    @torch.no_grad()
    def load_data(self, frame_ids=None, rgb_threshold=0.01):
        """video_dataparser.py:44-62: frames, flows, soft masks, flow ids, unique inverse (``self.unq_inv``)."""
        from . import flow_utils

        rgbs = self.load_video(frame_ids=frame_ids)
        frame_ids = list(frame_ids) if frame_ids is not None else list(range(rgbs.shape[0]))
        future_flows, past_flows, mask_bwds = self.load_flow(frame_ids, True, True, rgbs)
        flow_ids = flow_utils.get_flowid(rgbs, future_flows, mask_bwds, rgb_threshold=rgb_threshold)
        N, H, W = flow_ids.shape
        self.unq_inv = flow_utils.voxelization(flow_ids.view(-1, 1), id_range=N * H * W)
        return rgbs.permute(0, 2, 3, 1).reshape(-1, 3), None, None, future_flows, past_flows, mask_bwds


# ---- outputs (utils/VidToMe/utils.py:147-189) -----------------------------------------------------------------
def save_frames(frames: torch.Tensor, path: str, ext: str = "png", frame_ids=None) -> None:
    """One image file per frame, named by frame id (``0000.png`` ...)."""
    import cv2

    os.makedirs(path, exist_ok=True)
    ids = list(range(len(frames))) if frame_ids is None else list(frame_ids)
    for i, f in zip(ids, frames):
        img = (f.detach().float().clamp(0, 1) * 255).to(torch.uint8).permute(1, 2, 0).cpu().numpy()
        cv2.imwrite(os.path.join(path, "{:04}.{}".format(i, ext)), cv2.cvtColor(img, cv2.COLOR_RGB2BGR))


def save_video(frames: torch.Tensor, path: str, frame_ids=None, save_frame: bool = False, gif: bool = False, post_fix: str = "",
               fps: int = 30) -> str:
    """``output<post_fix>.mp4`` under ``path`` (OpenCV writer: torchvision.io.write_video / imageio are not available
    here, so the container is mp4v instead of the reference's libx264 and ``gif=True`` is not supported), plus the
    frames as PNGs when ``save_frame``.  Returns the video path."""
    import cv2

    if gif:
        raise TclError("save_video: GIF output needs imageio (not available); use gif=False")
    os.makedirs(path, exist_ok=True)
    ids = list(range(len(frames))) if frame_ids is None else list(frame_ids)
    sel = frames[ids]
    proc = (sel.detach().float().permute(0, 2, 3, 1) * 255).to(torch.uint8).cpu().numpy()       # truncation like the reference
    out = os.path.join(path, f"output{post_fix}.mp4")
    h, w = proc.shape[1:3]
    wr = cv2.VideoWriter(out, cv2.VideoWriter_fourcc(*"mp4v"), float(fps), (w, h))
    if not wr.isOpened():
        raise TclError(f"save_video: cannot open a writer for {out}")
    for img in proc:
        wr.write(cv2.cvtColor(img, cv2.COLOR_RGB2BGR))
    wr.release()
    print(f"[INFO] save video to {out}")
    if save_frame:
        save_frames(sel, os.path.join(path, f"frames{post_fix}"), frame_ids=ids)
    return out


def save_loss_curve(loss_list, output_path: str, title: str = "Loss Curve") -> str:
    """utils/general_utils.py: the reference plots with matplotlib (absent here); the values are written as text."""
    os.makedirs(output_path, exist_ok=True)
    p = os.path.join(output_path, f"{title}.txt")
    with open(p, "w") as f:
        f.write("\n".join(repr(float(v)) for v in loss_list))
    return p
