"""Convert the distributed CRIS checkpoint (``cris_best.pth``: ``{"state_dict": {"module.<key>": tensor}}`` as saved from
DistributedDataParallel) into the single-process state dict that ``CRIS(cris_pretrain=...)`` loads strictly
(``tunevlseg_b200/models/components/cris_model/__init__.py``; reference ``cris_model/__init__.py:64-69``).

Same command line and the same result as the reference's ``scripts/process_cris_checkpoint.py:5-25``.  Quirk kept on
purpose: the reference removes ``len(prefix) + 1`` characters from every key, so its default ``--prefix model.`` (6 + 1 =
7 characters) is what strips DDP's 7-character ``module.``; its validity check (``all(k for k in ...)``) can never fail -
here a key shorter than the cut raises, which is the only way that arithmetic can go wrong.
"""
from __future__ import annotations

import torch


def convert(state_dict: dict, prefix: str = "model.") -> dict:
    cut = len(prefix) + 1
    short = [k for k in state_dict if len(k) <= cut]
    if short:
        raise ValueError(f"Invalid checkpoint. All the keys of state_dict must start with `{prefix}` (too short: {short[:3]})")
    return {k[cut:]: v for k, v in state_dict.items()}


def main(checkpoint_input_path, checkpoint_output_path, prefix: str = "model.", pickle_protocol: int = 5) -> None:
    from ..models.components.cris_model import load_checkpoint

    checkpoint = load_checkpoint(checkpoint_input_path)
    torch.save(convert(checkpoint["state_dict"], prefix), checkpoint_output_path, pickle_protocol=pickle_protocol)


if __name__ == "__main__":
    from argparse import ArgumentParser

    parser = ArgumentParser(description="A script to convert the distributed checkpoint to a single machine checkpoint.")
    parser.add_argument("--checkpoint-input-path", type=str, default="pretrain/cris_best.pth", help="Path to the checkpoint to convert.")
    parser.add_argument("--checkpoint-output-path", type=str, default="pretrain/cris_best_single.pth", help="Path to save the converted checkpoint.")
    parser.add_argument("--prefix", type=str, default="model.", help="The prefix of the state_dict in the checkpoint.")
    parser.add_argument("--pickle-protocol", type=int, default=5, help="The protocol to use when pickling the checkpoint.")
    main(**vars(parser.parse_args()))
