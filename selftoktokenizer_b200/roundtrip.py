"""The reference's driver script (test.py) on this library:

    python -m selftoktokenizer_b200.roundtrip --yml-path configs/selftok_256_512tok.yml --pretrained tokenizer_512_ckpt.pth \\
        --sd3_pretrained <stable-diffusion-3-medium-diffusers snapshot> --data_size 256 --images test.jpg

images -> SelftokPipeline.encoding -> token.npy -> SelftokPipeline.decoding -> re_<b>_<size>.png.  `--synthetic` replaces both
checkpoints by the seeded synthetic ones of selftoktokenizer_b200.synth (no file is read; smoke runs on a box without weights).
`--prepack-cache DIR` keeps the packed device state between runs (start-up in seconds instead of a torch.load of 8 GB)."""
from __future__ import annotations

import argparse

import numpy as np
import torch

from . import SelftokPipeline, parse_args_from_yaml, synth
from .config import SelftokDims
from .preprocess import load_images, save_image


def main(argv=None) -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--yml-path", type=str, default="configs/selftok_256_512tok.yml")
    ap.add_argument("--pretrained", type=str, default=None)
    ap.add_argument("--sd3_pretrained", type=str, default=None)
    ap.add_argument("--data_size", type=int, default=256)
    ap.add_argument("--images", type=str, nargs="+", default=["./test.jpg"])
    ap.add_argument("--out-prefix", type=str, default="./re")
    ap.add_argument("--tokens", type=str, default="./token.npy")
    ap.add_argument("--synthetic", action="store_true")
    ap.add_argument("--prepack-cache", type=str, default=None)
    ap.add_argument("--device", type=str, default="cuda")
    args = ap.parse_args(argv)

    cfg = parse_args_from_yaml(args.yml_path)
    if args.synthetic:
        from .pipeline import DeviceVAE
        dims = SelftokDims.from_cfg(cfg, args.data_size)
        dev = torch.device(args.device if ":" in args.device else args.device + ":0")
        model = SelftokPipeline(cfg=cfg, ckpt_path=None, sd3_path=None, datasize=args.data_size, device=dev,
                                state_dict=synth.synth_state_dict(dims, device=dev), vae=DeviceVAE(synth.synth_vae_state_dict(ch=128, device=dev), dev))
    else:
        model = SelftokPipeline(cfg=cfg, ckpt_path=args.pretrained, sd3_path=args.sd3_pretrained, datasize=args.data_size, device=args.device,
                                prepack_cache=args.prepack_cache)
    images = load_images(args.images, args.data_size).to(args.device)
    tokens = model.encoding(images, device=args.device)
    np.save(args.tokens, tokens.detach().cpu().numpy())
    tokens = np.load(args.tokens)
    images = model.decoding(tokens, device=args.device)
    for b in range(len(images)):
        save_image(images[b], f"{args.out_prefix}_{b}_{args.data_size}.png")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
