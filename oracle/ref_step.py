"""TEST / BASELINE INFRASTRUCTURE ONLY (oracle/): run the REAL reference's own train step (`SinSKITGModel.set_input` +
`optimize_parameters`, models/sinskitG_model.py:601-793) on the host CPU through its own option parser, for bench.py's
`--impl reference` arm (`cpu_baseline.kind = "reference"`).  Only usable where the reference tree is mounted (the build
container); the GPU box has no /root/reference and falls back to the oracle port.  Nothing is copied: the reference is imported
in place (oracle/ref_loader.py supplies stubs for its absent third-party imports; LPIPS / vision-aided terms are switched off
with the reference's own flags)."""
import contextlib
import io
import os

from .ref_loader import REF_ROOT, load_reference, reference_available  # noqa: F401


def reference_step_runner(S, NT=64, NF=32, netG="resnet_9blocks", ngf=64, ndf=64, seed=0, ckpt_dir="/tmp/vts_ref_bench"):
    """-> (step() callable running one reference train step on a seeded synthetic batch, description string)."""
    import torch
    from . import skit_oracle as O
    load_reference()
    cwd = os.getcwd()
    os.chdir(REF_ROOT)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            from models import create_model
            from options.train_options import TrainOptions
            to = TrainOptions()
            to.cmd_line = ("--model sinskitG --gpu_ids -1 --name bench --checkpoints_dir %s --crop_size %d --center_w %d --center_h %d "
                           "--lambda_G1_lpips 0 --lambda_G2_lpips 0 --use_vision_aided_loss False --batch_size_G2 %d "
                           "--add_fake_T_sample_size %d --netG %s --ngf %d --ndf %d"
                           % (ckpt_dir, S, S, S, NT, NF, netG, ngf, ndf)).split()
            torch.manual_seed(seed)
            opt = to.parse()
            model = create_model(opt)
            model.setup(opt)
        model.train()
    finally:
        os.chdir(cwd)
    batch = O.synthetic_batch(S, NT=NT, seed=seed)

    def step():
        with contextlib.redirect_stdout(io.StringIO()):
            model.set_input(batch, phase="train")
            model.optimize_parameters(1)

    return step, "the unmodified reference's SinSKITGModel.optimize_parameters (imported from %s)" % REF_ROOT
