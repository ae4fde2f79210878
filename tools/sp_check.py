#!/usr/bin/env python
"""Multi-GPU parity of the Ulysses path: every rank runs the sequence-parallel forward; rank 0 also runs the same
model un-sharded and compares. Launch: torchrun --nproc-per-node N --master-addr 127.0.0.1 tools/sp_check.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from frameino_b200 import synth  # noqa: E402
from frameino_b200.ulysses import disable_sequence_parallel, enable_sequence_parallel  # noqa: E402
from frameino_b200.wan import WanTransformer3DModel  # noqa: E402


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if "cog" in sys.argv[1:]:
        run_cog_cases()
    else:
        run_cases()
    dist.destroy_process_group()


def run_cog_cases():
    """CogVideoX (joint text + video sequence, batch 2 = the pipeline's batched CFG), NCCL and fused peer exchange."""
    from frameino_b200.cogvideox import CogVideoXTransformer3DModel

    rank, world = dist.get_rank(), dist.get_world_size()
    cfg = dict(synth.COG_TINY)
    if world == 8:
        cfg["num_attention_heads"] = 8
    results = {}
    for name, (f, h, w), batch in [("default_canvas_b2", (3, 12, 16), 2), ("resized_canvas_b1", (3, 16, 12), 1)]:
        sd = synth.make_cog_state_dict(cfg, seed=0, dtype=torch.bfloat16)
        hidden, ts, text = synth.make_cog_inputs(cfg, f, h, w, n_id=1, batch=batch, dtype=torch.bfloat16)
        cos, sin = synth.cog_rope_tables(cfg["attention_head_dim"], h // 2, w // 2, f, 1, device="cuda")
        model = CogVideoXTransformer3DModel(**cfg)
        model.load_state_dict(sd)
        model = model.to_inference_dtype(torch.bfloat16).cuda().eval()
        args = dict(hidden_states=hidden.cuda(), encoder_hidden_states=text.cuda(), timestep=ts.cuda(),
                    image_rotary_emb=(cos, sin), return_dict=False)
        ref = model(**args)[0]
        for mode in ("nccl", "peer"):
            enable_sequence_parallel(model, mode=mode)
            out = model(**args)[0]
            out2 = model(**args)[0]
            assert torch.equal(out, out2), "sequence-parallel forward is not repeatable"
            disable_sequence_parallel(model)
            torch.cuda.synchronize()
            err = float((out.float() - ref.float()).abs().max() / ref.float().abs().max())
            errs = [None] * world
            dist.all_gather_object(errs, err)
            results[f"cog/{mode}/{name}"] = {"tokens": text.shape[1] + (f + 1) * (h // 2) * (w // 2), "batch": batch,
                                             "rel_err_per_rank": errs}
    if rank == 0:
        print("SP_CHECK " + json.dumps(results))
        ok = all(e <= 2e-2 for r in results.values() for e in r["rel_err_per_rank"])
        print("SP_CHECK_OK" if ok else "SP_CHECK_FAILED", flush=True)


def run_cases():
    """Needs an initialised NCCL process group (also called by tools/step_breakdown.py --sp-check)."""
    rank = dist.get_rank()
    world = dist.get_world_size()
    results = {}
    # heads must divide by world: WAN_SMALL has 4 heads (ok for 2, 4); an 8-head variant for world 8
    cfg = dict(synth.WAN_SMALL)
    if world == 8:
        cfg["num_attention_heads"] = 8
    cases = [(m, n, sh, pt) for m in ("peer", "nccl")
             for n, sh, pt in [("even", (3, 32, 32), True), ("ragged", (2, 18, 22), True), ("scalar_t", (3, 16, 16), False)]]
    for mode, name, shape, per_token in cases:
        sd = synth.make_wan_state_dict(cfg, seed=0, dtype=torch.bfloat16)
        hidden, ts, text = synth.make_wan_inputs(cfg, *shape, n_id=1, text_len=16, text_true_len=11,
                                                 per_token_timestep=per_token, dtype=torch.bfloat16)
        model = WanTransformer3DModel(**cfg)
        model.load_state_dict(sd)
        model = model.to_inference_dtype(torch.bfloat16).cuda().eval()
        args = dict(hidden_states=hidden.cuda(), timestep=ts.cuda(), encoder_hidden_states=text.cuda(), return_dict=False)
        ref = model(**args)[0]
        enable_sequence_parallel(model, mode=mode)
        out = model(**args)[0]
        out2 = model(**args)[0]  # second forward reuses the exchange buffers
        assert torch.equal(out, out2), "sequence-parallel forward is not repeatable"
        disable_sequence_parallel(model)
        torch.cuda.synchronize()
        err = float((out.float() - ref.float()).abs().max() / ref.float().abs().max())
        errs = [None] * world
        dist.all_gather_object(errs, err)
        results[f"{mode}/{name}"] = {"tokens": (shape[0] + 1) * (shape[1] // 2) * (shape[2] // 2), "rel_err_per_rank": errs}
    if rank == 0:
        print("SP_CHECK " + json.dumps(results))
        ok = all(e <= 2e-2 for r in results.values() for e in r["rel_err_per_rank"])
        print("SP_CHECK_OK" if ok else "SP_CHECK_FAILED", flush=True)


if __name__ == "__main__":
    main()
