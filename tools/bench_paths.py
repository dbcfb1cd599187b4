"""Secondary paths of SURVEY §8 on the GPU, one JSON line each (development / profiles aid; bench.py is the contract).

  python tools/bench_paths.py --what zencoder,shape,ct,pipeline,train,gen512 [--B 32] [--steps 10] [--warmup 3]
  train under torchrun:  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_paths.py --what train

zencoder  a8   style encode, images/s                        (config 3 encode half)
shape     a10  hair+face encode and decode-by-code, images/s
ct        a11  encoder -> generator -> predictor MLP chain, codes/s
pipeline       config 3: encode -> edit -> decode chain with every network call on the B200 path, images/s
train     a12  config 5: one train.py loop iteration (D sub-step + G sub-step + both Adam updates), steps/s;
               world size > 1 adds the flat gradient all-reduce (NCCL)
gen512         config 4 per-GPU share: generator forward at 512x512, images/s
backend   8f.1 config 3 through BackendB200 (batched Backend.set_input_img + output incl. blending), images/s
blend     8f.2 postprocess_blending (blend mask + Poisson solve) at 256x256, images/s
All timings: CUDA events on the launching stream, W warm-up + K timed iterations, inputs resident on the device.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ctrlhair_b200 import parallel, synth  # noqa: E402

HAIR = 13
ZENC_GFLOP, SHAPE_GFLOP = 42.4, 37.1   # SURVEY §8d per image
PEAK_TF = 1363.8


def timed(fn, steps, warmup, stream=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if stream is None:
        e0.record()
    else:
        e0.record(stream)
    for _ in range(steps):
        fn()
    if stream is None:
        e1.record()
    else:
        e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


QUIET = False   # bench.py imports these functions and folds their dicts into its single JSON line


def emit(d):
    if not QUIET:
        print(json.dumps(d), flush=True)
    return d


def bench_zencoder(a, sd):
    from ctrlhair_b200.zencoder import ZencoderB200
    z = ZencoderB200(max_batch=a.B).load_state_dict(sd)
    img, lab = synth.make_image(a.B, 256).cuda(), synth.make_labels(a.B, 256, "blocky").cuda()
    ms = timed(lambda: z(img, lab), a.steps, a.warmup)
    return emit({"path": "zencoder (a8)", "B": a.B, "ms": ms, "images_per_s": a.B / ms * 1e3,
          "tflops_algorithmic": ZENC_GFLOP * a.B / ms, "frac_of_sustained_peak": ZENC_GFLOP * a.B / ms / PEAK_TF})


def bench_shape(a):
    from ctrlhair_b200.shape import ShapeGeneratorB200
    s = ShapeGeneratorB200(max_batch=a.B).load_state_dict(synth.make_shape_state_dict())
    hair, face = synth.make_shape_inputs(a.B)
    hair, face = hair.cuda(), face.cuda()
    hc, fc = s.forward_hair_encoder(hair, testing=True), s.forward_face_encoder(face)
    ms_e = timed(lambda: (s.forward_hair_encoder(hair, testing=True), s.forward_face_encoder(face)), a.steps, a.warmup)
    ms_d = timed(lambda: s.forward_decode_by_code(hc, fc), a.steps, a.warmup)
    return emit({"path": "shape nets (a10)", "B": a.B, "encode_ms": ms_e, "decode_ms": ms_d,
          "images_per_s": a.B / (ms_e + ms_d) * 1e3, "tflops_algorithmic": SHAPE_GFLOP * a.B / (ms_e + ms_d),
          "weights_mb_fp16": 482, "note": "241 M parameters: weight-bandwidth-bound at small B"})


def bench_ct(a):
    from ctrlhair_b200 import color_texture as ct
    g_sd, d_sd, p_sd = synth.make_ct_state_dicts()
    G, D, P = (ct.EigenGeneratorB200().load_state_dict(g_sd), ct.CodeEncoderB200().load_state_dict(d_sd),
               ct.PredictorB200().load_state_dict(p_sd))
    code = synth.make_ct_inputs(a.B)["code"].cuda()

    def chain():
        pred = P({"code": code})
        return ct.edit_infer(D, G, code, {"rgb_mean": pred["rgb_mean"], "pca_std": pred["pca_std"]})
    ms = timed(chain, a.steps * 5, a.warmup)
    return emit({"path": "colour/texture MLPs (a11): predictor + encoder + generator", "B": a.B, "ms": ms,
          "codes_per_s": a.B / ms * 1e3, "launches": 3, "bound": "launch latency (0.93 M parameters)"})


def bench_pipeline(a, sd):
    """Config 3: Backend.parse_img + Backend.output network calls (ui/backend.py:67-106,147-175), batch B."""
    from ctrlhair_b200 import color_texture as ct
    from ctrlhair_b200.generator import SeanGeneratorB200
    from ctrlhair_b200.shape import ShapeGeneratorB200
    from ctrlhair_b200.zencoder import ZencoderB200
    B = a.B
    shp = ShapeGeneratorB200(max_batch=B).load_state_dict(synth.make_shape_state_dict())
    zen = ZencoderB200(max_batch=B).load_state_dict(sd)
    gen = SeanGeneratorB200(max_batch=B).load_state_dict(sd)
    g_sd, d_sd, p_sd = synth.make_ct_state_dicts()
    G, D, P = (ct.EigenGeneratorB200().load_state_dict(g_sd), ct.CodeEncoderB200().load_state_dict(d_sd),
               ct.PredictorB200().load_state_dict(p_sd))
    labels_h = synth.make_labels(B, 256, "blocky", seed=99).pin_memory()
    img_h = synth.make_image(B, 256).pin_memory()
    out_h = torch.empty((B, 3, 256, 256)).pin_memory()

    def chain():
        labels = labels_h.cuda(non_blocking=True)
        img = img_h.cuda(non_blocking=True)
        oh = torch.zeros((B, 19, 256, 256), device="cuda").scatter_(1, labels[:, None].long(), 1.0)
        hair, face = oh[:, [HAIR]], torch.cat([oh[:, :HAIR], oh[:, HAIR + 1:]], 1)
        hc, fc = shp.forward_hair_encoder(hair, testing=True), shp.forward_face_encoder(face)
        lab = shp.forward_decode_by_code(hc, fc).argmax(1).to(torch.uint8)
        codes = zen(img, labels)
        hair_code = codes[:, HAIR].contiguous()
        pred = P({"code": hair_code})
        feat = ct.edit_infer(D, G, hair_code, {"rgb_mean": pred["rgb_mean"] * 0.5 + 0.2, "pca_std": pred["pca_std"]})
        codes[:, HAIR] = feat
        out = gen.forward_labels(lab, codes, seed=1)
        out_h.copy_(out, non_blocking=True)
    ms = timed(chain, a.steps, a.warmup)
    return emit({"path": "config 3: Backend encode -> edit -> decode (network calls only; parsing/blending out of scope)",
          "B": B, "ms": ms, "images_per_s": B / ms * 1e3, "h2d_bytes": int(labels_h.numel() + img_h.numel() * 4),
          "d2h_bytes": int(out_h.numel() * 4)})


def bench_train(a):
    from ctrlhair_b200 import ct_train
    own_pg = not (torch.distributed.is_available() and torch.distributed.is_initialized())
    rank, local_rank, world = parallel.init_process_group()
    torch.cuda.set_device(local_rank)
    B = a.train_batch
    s = ct_train.SolverB200(None, "cuda:%d" % local_rank, batch_size=B, use_graph=getattr(a, "train_mode", "graph"))
    s.load_state_dicts(*synth.make_ct_train_state_dicts())
    batches = [synth.make_ct_train_batch(B, 5000 + 100 * rank + i) for i in range(4)]   # rank-distinct data
    batches = [{k: v.cuda() for k, v in b.items()} for b in batches]
    it = [0]

    def step():
        for i in range(2):
            ld = {}
            s.forward(batches[(it[0] * 2 + i) % 4])
            if i == 0:
                s.forward_d(ld)
                ct_train.train(s.cfg, ld, optimizers=[s.D_optimizer])
            else:
                s.forward_g(ld)
                ct_train.train(s.cfg, ld, optimizers=[s.G_optimizer])
        it[0] += 1
    for _ in range(a.warmup):
        step()
    s.synchronize()
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s.stream)
    n = a.steps * 5
    for _ in range(n):
        step()
    e1.record(s.stream)
    s.synchronize()
    ms = e0.elapsed_time(e1) / n
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t)
    finite = bool(torch.isfinite(s.losses).all())
    res = None
    if rank == 0:
        res = emit({"path": "config 5: colour/texture train.py iteration (D + G sub-steps, 2 Adam updates)", "n_gpus": world,
              "batch_per_gpu": B, "global_batch": B * world, "ms_per_step": ms, "steps_per_s": 1e3 / ms,
              "samples_per_s": B * world / ms * 1e3, "launches_per_step": s.launches(0) + s.launches(1) + 2,
              "mode": {2: "persistent cooperative kernel per sub-step", 1: "explicit dependency CUDA graph per sub-step", 0: "plain launches"}[s.mode],
              "operations_and_%s" % ("barriers" if s.mode == 2 else "graph_edges"):
                  [s.schedule(0), s.schedule(1)] if s.mode else None,
              "allreduce_per_step": 2 if world > 1 else 0, "dtype": "f32", "losses_finite": finite})
    if world > 1 and own_pg:
        torch.distributed.destroy_process_group()
    return res


def bench_blend(a):
    """HairEditor.postprocess_blending (hair_editor.py:257-308) for B images, device resident.
    (The CPU figure beside it comes from tests/_cpu_baselines.py, which times the oracle port: nothing outside tests/,
    smoke() and bench.py's cpu_baseline leg touches oracle/.)"""
    import numpy as np
    from ctrlhair_b200 import blend
    B = a.B
    cases = [synth.make_blend_case(256, 256, 900 + i) for i in range(B)]
    face = torch.from_numpy(np.stack([c[0] for c in cases])).cuda()
    res = torch.from_numpy(np.stack([c[1] for c in cases])).cuda().permute(0, 3, 1, 2).float().div(127.5).sub(1).contiguous()
    fp = torch.from_numpy(np.stack([c[2] for c in cases])).cuda()
    tp = torch.from_numpy(np.stack([c[3] for c in cases])).cuda()
    ms = timed(lambda: blend.postprocess_blending(face, res, fp, tp), a.steps, a.warmup)
    src_u8 = blend.image_to_u8(res)
    mask = 1 - blend.blend_mask(tp, fp)
    _, stats = blend.poisson_blending(face, src_u8, mask, return_stats=True)
    ms_solve = timed(lambda: blend.poisson_blending(face, src_u8, mask), a.steps, a.warmup)
    it = float(stats[..., 0].mean())
    return emit({"path": "8f.2 postprocess_blending 256x256 (mask + fp64 CG Poisson solve, 3 channels)", "B": B, "ms": ms,
          "images_per_s": B / ms * 1e3, "poisson_kernel_ms": ms_solve, "cg_iterations_mean": it,
          "cg_iterations_max": float(stats[..., 0].max()), "residual_max": float(stats[..., 1].max()),
          "us_per_iteration_per_wave": ms_solve * 1e3 / it / max(1.0, B * 3 / 16.0),
          "unknown_fraction": float(torch.as_tensor(mask).float().mean())})


def bench_backend(a, sd):
    """Config 3 through BackendB200: set_input_img (shape encode/decode, style encode, colour predictor, RGB->HSV, code
    encoder) + output (HSV->RGB, feature generator, generator, blend mask + Poisson blending), host buffers in and out,
    nothing but the parsing network (BiSeNet, out of scope) missing from ui/backend.py's chain."""
    import numpy as np
    from ctrlhair_b200.backend import BackendB200
    B = a.B
    cases = [synth.make_blend_case(256, 256, 700 + i) for i in range(B)]
    img_h = torch.from_numpy(np.stack([c[0] for c in cases])).pin_memory()
    lab_h = torch.from_numpy(np.stack([c[2] for c in cases])).pin_memory()
    out_h = torch.empty((B, 256, 256, 3), dtype=torch.uint8).pin_memory()
    be = BackendB200(sd, synth.make_shape_state_dict(), synth.make_ct_state_dicts(),
                     median_codes=synth.make_codes(1, seed=4321)[0], max_batch=B, blending=True)

    def chain():
        be.set_input_img(img_h.cuda(non_blocking=True), lab_h.cuda(non_blocking=True))
        out_h.copy_(be.output(), non_blocking=True)
    ms = timed(chain, a.steps, a.warmup)
    be.blending = False
    ms_nb = timed(chain, a.steps, a.warmup)
    return emit({"path": "config 3 via BackendB200: set_input_img + output incl. Poisson blending (parsing network excluded)",
          "B": B, "ms": ms, "images_per_s": B / ms * 1e3, "ms_without_blending": ms_nb,
          "images_per_s_without_blending": B / ms_nb * 1e3,
          "h2d_bytes": int(img_h.numel() + lab_h.numel()), "d2h_bytes": int(out_h.numel())})


def load_example_faces(n):
    """The reference's 50 example faces (imgs/*.png, all 256x256x3; staged by baseline/make_ref.py), RGB uint8 [n,256,256,3]
    (cycled when n > 50).  None when the staged copy is absent."""
    import glob
    import numpy as np
    d = os.path.join(ROOT, "baseline", "_ref", "imgs")
    files = sorted(glob.glob(os.path.join(d, "*.png")))
    if not files:
        return None, 0
    import cv2
    imgs = [cv2.cvtColor(cv2.imread(f, cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB) for f in files]
    imgs = [cv2.resize(im, (256, 256)) if im.shape[:2] != (256, 256) else im for im in imgs]   # ui/backend.py:69
    return np.stack([imgs[i % len(imgs)] for i in range(n)]), len(files)


def bench_config3(a, sd):
    """BASELINE.json config 3 as stated: Backend encode -> edit -> decode on imgs/*.png, batch 32, one B200 — every
    network of ui/backend.py:67-106,147-175 on the GPU path INCLUDING the face parser (BiSeNet, my_parsing_util.py:31-47),
    host uint8 images in, host uint8 images out.  Weights are the synthetic checkpoints (no trained ones ship with the
    reference), so the masks are what the synthetic parser makes of the faces; the work per image is the same."""
    from ctrlhair_b200.backend import BackendB200
    B = a.B
    faces, nfiles = load_example_faces(B)
    if faces is None:
        return emit({"path": "config 3 on imgs/*.png", "unavailable": "baseline/_ref/imgs not staged (baseline/make_ref.py)"})
    img_h = torch.from_numpy(faces).pin_memory()
    out_h = torch.empty((B, 256, 256, 3), dtype=torch.uint8).pin_memory()
    be = BackendB200(sd, synth.make_shape_state_dict(), synth.make_ct_state_dicts(),
                     median_codes=synth.make_codes(1, seed=4321)[0], max_batch=B, blending=True,
                     parsing_sd=synth.make_bisenet_state_dict())

    def chain():
        be.set_input_img(img_h)                      # parses (PIL-exact resize + BiSeNet, both on the GPU), encodes
        out_h.copy_(be.output(), non_blocking=True)  # edits, decodes, blends
        torch.cuda.current_stream().synchronize()
    import time as _t
    for _ in range(a.warmup):
        chain()
    t0 = _t.perf_counter()
    for _ in range(a.steps):
        chain()
    ms = (_t.perf_counter() - t0) / a.steps * 1e3
    # where the time goes: the parser alone (resize + network on the device), and what the reference's host-side PIL
    # resize of the same 32 images costs (it is NOT on this path any more: bisenet.resize_bilinear_u8 is bit exact)
    img_d = img_h.cuda()
    ms_parse = timed(lambda: be.face_parser.get_mask_device(img_d, 256), a.steps, a.warmup)
    t0 = _t.perf_counter()
    for im in faces:
        be.face_parser.resize_to_network(im, 512)
    ms_resize = (_t.perf_counter() - t0) * 1e3
    be.blending = False
    mask = be.get_mask(img_h)
    ms_nb = timed(lambda: (be.set_input_img(img_h.cuda(non_blocking=True), mask), be.output()), a.steps, a.warmup)
    return emit({"path": "config 3: Backend encode -> edit -> decode on the reference's imgs/*.png (parser included)",
                 "B": B, "png_files": nfiles, "ms": ms, "images_per_s": B / ms * 1e3, "timing": "host wall clock, host buffers in/out",
                 "face_parser_gpu_ms": ms_parse, "reference_host_pil_resize_ms_not_on_path": ms_resize,
                 "ms_without_parser_and_blending_device_timed": ms_nb,
                 "h2d_bytes": int(img_h.numel()), "d2h_bytes": int(out_h.numel()),
                 "weights": "synthetic checkpoints (reference ships none); masks from the synthetic BiSeNet"})


def bench_gen512(a, sd):
    from ctrlhair_b200 import flops as flopmodel
    from ctrlhair_b200.generator import SeanGeneratorB200
    B = a.B512
    gen = SeanGeneratorB200(crop=512, max_batch=B).load_state_dict(sd)
    labels, codes = synth.make_labels(B, 512, "blocky").cuda(), synth.make_codes(B).cuda()
    out = torch.empty((B, 3, 512, 512), device="cuda")
    ms = timed(lambda: gen.forward_labels(labels, codes, seed=1, out=out), max(2, a.steps // 2), a.warmup)
    _, fact = flopmodel.generator_macs(512)
    return emit({"path": "config 4 per-GPU share: generator forward 512x512", "B": B, "ms": ms, "images_per_s": B / ms * 1e3,
          "tflops_algorithmic": 2 * fact * B / ms / 1e9, "frac_of_sustained_peak": 2 * fact * B / ms / 1e9 / PEAK_TF,
          "finite": bool(torch.isfinite(out).all())})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="zencoder,shape,ct,pipeline,train")
    ap.add_argument("--B", type=int, default=32)
    ap.add_argument("--B512", type=int, default=16)
    ap.add_argument("--train-batch", type=int, default=32)
    ap.add_argument("--train-mode", default="graph", help="graph | persistent")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    what = a.what.split(",")
    sd = synth.make_state_dict() if any(w in what for w in ("zencoder", "pipeline", "gen512", "backend", "config3")) else None
    if "zencoder" in what:
        bench_zencoder(a, sd)
    if "shape" in what:
        bench_shape(a)
    if "ct" in what:
        bench_ct(a)
    if "pipeline" in what:
        bench_pipeline(a, sd)
    if "gen512" in what:
        bench_gen512(a, sd)
    if "config3" in what:
        bench_config3(a, sd)
    if "blend" in what:
        bench_blend(a)
    if "backend" in what:
        bench_backend(a, sd)
    if "train" in what:
        bench_train(a)


if __name__ == "__main__":
    main()
