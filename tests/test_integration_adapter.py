"""The reference-side binding (integration/b200sinskitG_model.py, b200skitG_model.py): the reference's own model discovery
(models/__init__.py:24-67) finds the adapter classes and its own option parser (options/base_options.py:221-258) parses
`--model b200sinskitG` with every sinskitG option plus the B200 extras.  Runs only where the reference tree is present (the
build container); the step itself is covered on the GPU by tests/test_basemodel_contract_gpu.py."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
from oracle import ref_loader  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref_models():
    ref_loader.load_reference()
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        import models
    integ = os.path.join(ROOT, "integration")
    if integ not in models.__path__:
        models.__path__.append(integ)       # stands for copying the two files into the reference's models/ directory
    return models


@pytest.mark.parametrize("name,cls_name", [("b200sinskitG", "SinSKITGModel"), ("b200skitG", "SKITGModel")])
def test_reference_discovers_the_adapter(ref_models, name, cls_name):
    import vts_b200
    from models.base_model import BaseModel
    cls = ref_models.find_model_using_name(name)
    assert issubclass(cls, BaseModel) and issubclass(cls, getattr(vts_b200, cls_name))
    # the B200 implementations win the method resolution for everything train.py / test.py call
    for meth in ("set_input", "forward", "optimize_parameters", "test", "setup", "parallelize", "train", "eval", "get_current_losses",
                 "get_current_visuals", "get_current_metrics", "get_image_paths", "update_learning_rate", "save_networks", "load_networks"):
        assert getattr(cls, meth).__qualname__.split(".")[0] in ("SinSKITGModel", "SKITGModel"), meth
        assert getattr(cls, meth) is not getattr(BaseModel, meth, None), meth
    assert not getattr(cls, "__abstractmethods__", None)


def test_reference_option_parser_accepts_the_adapter(ref_models, tmp_path):
    import contextlib
    import io
    cwd = os.getcwd()
    os.chdir(ref_loader.REF_ROOT)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            from options.train_options import TrainOptions
            to = TrainOptions()
            to.cmd_line = ("--model b200sinskitG --gpu_ids -1 --name adapter --checkpoints_dir %s --netG resnet_9blocks --ngf 64 "
                           "--lambda_NCE 1.0 --cuda_graph false" % tmp_path).split()
            opt = to.parse()
            ts = TrainOptions()
            ts.cmd_line = ("--model sinskitG --gpu_ids -1 --name plain --checkpoints_dir %s" % tmp_path).split()
            ref = ts.parse()
    finally:
        os.chdir(cwd)
    assert opt.model == "b200sinskitG" and opt.lambda_NCE == 1.0 and opt.cuda_graph is False and opt.num_patches == 256
    assert opt.use_vision_aided_loss is False and opt.lambda_G1_lpips == 0.0
    # every option of the reference model is present with the reference's default (bar the three third-party switches)
    skip = {"model", "name", "netG", "ngf", "use_vision_aided_loss", "lambda_G1_lpips", "lambda_G2_lpips"}
    for k, v in vars(ref).items():
        assert hasattr(opt, k), k
        if k not in skip:
            assert getattr(opt, k) == v, k
    # and everything the B200 model reads without a fallback is there
    import vts_b200
    for k in vars(vts_b200.default_options()):
        if k not in ("cuda_graph", "cuda_graph_warmup", "run_full_res_D2", "lpips_state", "allow_random_lpips", "isTrain", "gpu_ids",
                     # older spellings default_options() keeps beside the reference's names; the model falls back between them
                     "num_D", "input_nc", "output_nc",
                     # skitG-only options (models/skitG_model.py:255-276), read with their defaults when absent
                     "use_style_code", "style_code_mode", "style_code_mapping_mode", "style_code_dim", "num_layer_style_code"):
            assert hasattr(opt, k), k


def test_reference_discovers_the_dataset_adapter(ref_models):
    """data/__init__.py:18-40 finds `b200singleskit_dataset.py`, and its option setter carries the reference dataset's own flags."""
    import argparse
    import data
    import vts_b200
    from data.base_dataset import BaseDataset
    from data.singleskit_dataset import SingleSkitDataset as RefDataset
    integ = os.path.join(ROOT, "integration")
    if integ not in data.__path__:
        data.__path__.append(integ)
    multi = data.find_dataset_using_name("b200skit")
    assert issubclass(multi, BaseDataset) and issubclass(multi, vts_b200.SkitDataset) and not getattr(multi, "__abstractmethods__", None)
    cls = data.find_dataset_using_name("b200singleskit")
    assert issubclass(cls, BaseDataset) and issubclass(cls, vts_b200.SingleSkitDataset)
    for meth in ("__getitem__", "__len__", "preprocess_data", "find_validate_touch_patches_and_coords"):
        assert getattr(cls, meth).__qualname__.split(".")[0] == "SingleSkitDataset", meth
    assert not getattr(cls, "__abstractmethods__", None)
    for is_train in (True, False):
        mine = data.get_option_setter("b200singleskit")(argparse.ArgumentParser(), is_train).parse_args([])
        ref = RefDataset.modify_commandline_options(argparse.ArgumentParser(), is_train).parse_args([])
        assert vars(mine) == vars(ref)
