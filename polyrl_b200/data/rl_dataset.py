"""Prompt dataset for RL: parquet / jsonl rows -> chat template -> tokenise -> left-pad / truncate, with
over-long filtering and resumable sampling (reference: 3rdparty/verl/verl/utils/dataset/rl_dataset.py:69-463).
A whitespace/byte tokenizer is provided for offline tests (the reference's digit-completion env plays the
same role, tests/special_e2e/envs/digit_completion/)."""
from __future__ import annotations

import json
from pathlib import Path
from typing import Any

import numpy as np
import torch
from torch.utils.data import Dataset, Sampler


class ByteTokenizer:
    """Offline tokenizer: UTF-8 bytes + specials; enough for plumbing tests and synthetic runs."""

    pad_token_id, eos_token_id, bos_token_id = 0, 1, 2
    vocab_size = 259

    def encode(self, text: str, add_special_tokens: bool = False) -> list[int]:
        ids = [b + 3 for b in text.encode("utf-8")]
        return ([self.bos_token_id] + ids) if add_special_tokens else ids

    def decode(self, ids, skip_special_tokens: bool = True) -> str:
        bs = bytes(int(i) - 3 for i in ids if 3 <= int(i) < 259)      # ids above the byte range (larger model vocab) are dropped
        return bs.decode("utf-8", errors="replace")

    def apply_chat_template(self, messages: list[dict], add_generation_prompt: bool = True, tokenize: bool = False, **_kw):
        text = "".join(f"<{m['role']}>{m['content']}\n" for m in messages) + ("<assistant>" if add_generation_prompt else "")
        return self.encode(text) if tokenize else text


def load_tokenizer(path: str | None):
    if not path or path == "byte":
        return ByteTokenizer()
    from transformers import AutoTokenizer

    tok = AutoTokenizer.from_pretrained(path)
    if tok.pad_token_id is None:
        tok.pad_token_id = tok.eos_token_id
    return tok


def _read_rows(files) -> list[dict]:
    if isinstance(files, (str, Path)):
        files = [files]
    rows: list[dict] = []
    for f in files:
        f = str(f)
        if f.endswith(".parquet"):
            import pyarrow.parquet as pq

            rows += pq.read_table(f).to_pylist()
        elif f.endswith(".jsonl"):
            rows += [json.loads(l) for l in Path(f).read_text().splitlines() if l.strip()]
        elif f.endswith(".json"):
            rows += json.loads(Path(f).read_text())
        else:
            raise ValueError(f"unsupported data file {f}")
    return rows


def left_pad(ids: list[int], max_len: int, pad_id: int, truncation: str = "error"):
    if len(ids) > max_len:
        if truncation == "left":
            ids = ids[-max_len:]
        elif truncation == "right":
            ids = ids[:max_len]
        elif truncation == "middle":
            half = max_len // 2
            ids = ids[:half] + ids[-(max_len - half):]
        else:
            raise ValueError(f"prompt of {len(ids)} tokens exceeds max_prompt_length={max_len} (truncation='error')")
    n = len(ids)
    input_ids = torch.full((max_len,), pad_id, dtype=torch.long)
    input_ids[max_len - n:] = torch.tensor(ids, dtype=torch.long)
    mask = torch.zeros(max_len, dtype=torch.long)
    mask[max_len - n:] = 1
    pos = (mask.cumsum(0) - 1).clamp_min(0)
    return input_ids, mask, pos, ids


def load_image(obj):
    """Dataset image cell -> PIL image: a PIL image, ``{"bytes": ...}`` / ``{"path": ...}`` (HF datasets / parquet encoding), a file
    path, or an [H, W, C] uint8 array (reference: verl/utils/dataset/vision_utils.py:25-60 process_image)."""
    import io

    from PIL import Image

    if isinstance(obj, Image.Image):
        return obj
    if isinstance(obj, dict):
        if obj.get("bytes") is not None:
            return Image.open(io.BytesIO(obj["bytes"]))
        if obj.get("path"):
            return Image.open(obj["path"])
        if obj.get("image") is not None:
            return load_image(obj["image"])
        raise ValueError(f"image cell without bytes / path: {list(obj)}")
    if isinstance(obj, (bytes, bytearray)):
        return Image.open(io.BytesIO(obj))
    if isinstance(obj, (str, Path)):
        return Image.open(obj)
    return Image.fromarray(np.asarray(obj, dtype=np.uint8))


def load_video(obj, fps: float = 2.0, nframes: int | None = None, min_frames: int = 4, max_frames: int = 768,
               frame_factor: int = 2) -> tuple[list, float]:
    """Dataset video cell -> (frames as PIL images, sampling rate in frames / second).

    Accepts a frame list (passed through), ``{"frames": [...], "fps": f}``, or a video FILE — a path, ``{"path" | "video": path,
    "fps": f, "nframes": n, "video_start": s, "video_end": s}`` — decoded with OpenCV.  Frames are sampled uniformly: ``nframes``
    if given, else ``fps`` frames per second of footage clamped to [min_frames, max_frames], and the count is rounded to a multiple
    of ``frame_factor`` (the tower's temporal patch) like the reference's ``fetch_video`` (verl/utils/dataset/vision_utils.py:
    63-125 -> qwen_vl_utils)."""
    if isinstance(obj, (list, tuple, np.ndarray)) and not isinstance(obj, (str, bytes)):
        return [load_image(f) for f in obj], float(fps)
    spec = dict(obj) if isinstance(obj, dict) else {"path": obj}
    if spec.get("frames") is not None:
        return [load_image(f) for f in spec["frames"]], float(spec.get("fps", fps))
    path = spec.get("path") or spec.get("video")
    if not path:
        raise ValueError(f"video cell without frames / path: {list(spec)}")
    try:
        import cv2
    except ImportError as e:                                   # pragma: no cover
        raise RuntimeError("decoding video files needs OpenCV (cv2); pass the video as a list of frames instead") from e
    from PIL import Image

    cap = cv2.VideoCapture(str(path))
    if not cap.isOpened():
        raise FileNotFoundError(f"cannot open video {path!r}")
    total = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
    src_fps = float(cap.get(cv2.CAP_PROP_FPS)) or 30.0
    first = max(0, int(round(float(spec.get("video_start", 0.0)) * src_fps)))
    last = min(total - 1, int(round(float(spec["video_end"]) * src_fps))) if spec.get("video_end") is not None else total - 1
    span = max(1, last - first + 1)
    want_fps = float(spec.get("fps", fps))
    n = spec.get("nframes", nframes)
    if n is None:
        n = span / src_fps * want_fps
        n = min(max(n, min_frames), max_frames, span)
    n = max(frame_factor, int(round(float(n) / frame_factor)) * frame_factor)
    n = min(n, max(frame_factor, span // frame_factor * frame_factor))
    idx = np.linspace(first, last, n).round().astype(int).tolist()
    frames, pos = [], -1
    for i in idx:
        if i != pos + 1:
            cap.set(cv2.CAP_PROP_POS_FRAMES, i)
        ok, bgr = cap.read()
        pos = i
        if not ok:                                             # container reported more frames than it holds: repeat the last one
            if not frames:
                raise ValueError(f"no decodable frame in {path!r}")
            frames.append(frames[-1])
            continue
        frames.append(Image.fromarray(cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB)))
    cap.release()
    return frames, n / (span / src_fps)


class RLHFDataset(Dataset):
    """``model_cfg`` with a vision tower turns on the multi-modal path: rows may carry ``image_key`` (list of images) and mark their
    place in the prompt with ``<image>``; items then hold ``multi_modal_inputs`` (patches + grids), prompt ids with one placeholder
    token per merged patch and 3-axis ``position_ids`` [3, L] (reference: rl_dataset.py:181-230, 299-405)."""

    IMAGE_MARK = "<image>"
    VIDEO_MARK = "<video>"

    def __init__(self, data_files, tokenizer, prompt_key: str = "prompt", max_prompt_length: int = 512,
                 filter_overlong_prompts: bool = False, truncation: str = "error", return_raw_chat: bool = False,
                 model_cfg=None, image_key: str = "images", min_pixels: int | None = None, max_pixels: int | None = None,
                 video_key: str = "videos", apply_chat_template_kwargs: dict | None = None):
        self.chat_kwargs = dict(apply_chat_template_kwargs or {})       # e.g. {"reasoning_effort": "medium", "enable_thinking": False}
        self.tok, self.key, self.max_len = tokenizer, prompt_key, max_prompt_length
        self.truncation, self.return_raw_chat = truncation, return_raw_chat
        self.vl_cfg = model_cfg if (model_cfg is not None and getattr(model_cfg, "vision", None) is not None) else None
        self.image_key, self.min_pixels, self.max_pixels = image_key, min_pixels, max_pixels
        self.video_key = video_key
        self.rows = _read_rows(data_files)
        if filter_overlong_prompts:
            self.rows = [r for r in self.rows if self._prompt_len(r) <= max_prompt_length]

    def _prompt_len(self, row: dict) -> int:
        if self._has_vision(row):
            return len(self._tokenize_mm(row)[0])
        return len(self._tokenize(row))

    def _has_vision(self, row: dict) -> bool:
        return self.vl_cfg is not None and bool(row.get(self.image_key) or row.get(self.video_key))

    def _chat_text(self, row: dict) -> str:
        p = row[self.key]
        if isinstance(p, (list, np.ndarray)) and len(p) and isinstance(p[0], dict):
            return self.tok.apply_chat_template(list(p), add_generation_prompt=True, tokenize=False, **self.chat_kwargs)
        return str(p)

    def _tokenize_mm(self, row: dict):
        """-> (prompt ids with expanded placeholders, pixel_values [N, patch_dim], grids, seconds per temporal patch of the videos).
        ``<image>`` / ``<video>`` marks in the prompt consume the row's images / videos in order; a video is a list of frames,
        ``{"frames": [...], "fps": f}``, or a video file (path / ``{"path": ..., "fps": ..., "nframes": ...}``, see ``load_video``)."""
        import re

        from ..models import vision as V

        cfg = self.vl_cfg
        images = [load_image(im) for im in (row.get(self.image_key) or [])]
        videos = list(row.get(self.video_key) or [])
        parts = re.split(f"({re.escape(self.IMAGE_MARK)}|{re.escape(self.VIDEO_MARK)})", self._chat_text(row))
        n_img, n_vid = parts.count(self.IMAGE_MARK), parts.count(self.VIDEO_MARK)
        if n_img != len(images) or n_vid != len(videos):
            raise ValueError(f"prompt has {n_img} {self.IMAGE_MARK} / {n_vid} {self.VIDEO_MARK} marks for {len(images)} images / {len(videos)} videos")
        it_img, it_vid = iter(images), iter(videos)
        ids: list[int] = []
        pix, grids, secs = [], [], []
        tp = cfg.vision.get("temporal_patch_size", 2)
        for seg in parts:
            if seg == self.IMAGE_MARK:
                pv, g = V.preprocess_image(cfg, next(it_img), self.min_pixels, self.max_pixels)
                tok = cfg.image_token_id
            elif seg == self.VIDEO_MARK:
                vid = next(it_vid)
                frames, fps = load_video(vid, frame_factor=tp)         # frame lists pass through, files are decoded (OpenCV)
                pv, g = V.preprocess_video(cfg, frames, self.min_pixels, self.max_pixels)
                secs.append(tp / fps)
                tok = cfg.video_token_id
            else:
                ids += list(self.tok.encode(seg, add_special_tokens=False)) if seg else []
                continue
            pix.append(pv)
            grids.append(g)
            ids += [cfg.vision_start_token_id] + [tok] * V.num_image_tokens(cfg, [g])[0] + [cfg.vision_end_token_id]
        return ids, torch.cat(pix, dim=0), grids, secs

    def _tokenize(self, row: dict) -> list[int]:
        return list(self.tok.encode(self._chat_text(row), add_special_tokens=False))

    def __len__(self) -> int:
        return len(self.rows)

    def __getitem__(self, i: int) -> dict[str, Any]:
        row = dict(self.rows[i])
        mm = None
        if self._has_vision(row):
            from ..models.vision import mrope_positions

            toks, pixel_values, grids, secs = self._tokenize_mm(row)
            ids, mask, pos, raw = left_pad(toks, self.max_len, self.tok.pad_token_id, "error")     # image runs cannot be cut
            pos3, _ = mrope_positions(self.vl_cfg, raw, grids, secs or None)
            pos = torch.zeros(3, self.max_len, dtype=torch.long)
            pos[:, self.max_len - len(raw):] = pos3
            # one grid list for images and videos, in prompt order (the placeholder token id tells them apart)
            mm = {"pixel_values": pixel_values, "image_grid_thw": torch.tensor(grids, dtype=torch.long)}
            if secs:
                mm["second_per_grid_ts"] = secs
        else:
            ids, mask, pos, raw = left_pad(self._tokenize(row), self.max_len, self.tok.pad_token_id, self.truncation)
            if self.vl_cfg is not None:
                pos = pos.unsqueeze(0).expand(3, -1).contiguous()          # text rows of a VLM dataset: same value on every axis
        out = {"input_ids": ids, "attention_mask": mask, "position_ids": pos, "raw_prompt_ids": raw,
               "data_source": row.get("data_source", "unknown"), "reward_model": row.get("reward_model", {"ground_truth": ""}),
               "extra_info": row.get("extra_info", {}), "index": (row.get("extra_info") or {}).get("index", i)}
        if self.return_raw_chat:
            out["raw_prompt"] = row[self.key]
        if self.vl_cfg is not None:
            out["multi_modal_inputs"] = mm
        # multi-turn rollouts: per-sample tool / interaction arguments travel in ``extra_info`` (reference:
        # verl/utils/dataset/rl_dataset.py:425-445 tools_kwargs / interaction_kwargs, need_tools_kwargs)
        ei = row.get("extra_info") or {}
        if ei.get("tools_kwargs") is not None:
            out["tools_kwargs"] = ei["tools_kwargs"]
        if ei.get("interaction_kwargs") is not None:
            out["interaction_kwargs"] = ei["interaction_kwargs"]
        return out


class ResumableRandomSampler(Sampler):
    """Seeded shuffling with a checkpointable cursor (plays the role of StatefulDataLoader state,
    reference ray_trainer.py:346-426,777-835)."""

    def __init__(self, n: int, seed: int = 1, shuffle: bool = True):
        self.n, self.seed, self.shuffle = n, seed, shuffle
        self.epoch, self.cursor = 0, 0

    def __iter__(self):
        while self.cursor < self.n:
            order = self._order()
            i = order[self.cursor]
            self.cursor += 1
            yield int(i)
        self.cursor, self.epoch = 0, self.epoch + 1

    def _order(self):
        if not self.shuffle:
            return np.arange(self.n)
        return np.random.default_rng(self.seed + self.epoch).permutation(self.n)

    def __len__(self) -> int:
        return self.n

    def state_dict(self) -> dict:
        return {"epoch": self.epoch, "cursor": self.cursor, "seed": self.seed}

    def load_state_dict(self, sd: dict) -> None:
        self.epoch, self.cursor, self.seed = sd["epoch"], sd["cursor"], sd["seed"]


def gsm8k_prompt(question: str) -> list[dict]:
    """GSM8K prompt template used by the reference's preprocessing (examples/data_preprocess/gsm8k.py:57-70)."""
    return [{"role": "user", "content": question + " Let's think step by step and output the final answer after \"####\"."}]
