"""`parse_config()` with the reference's flags, defaults and precedence (config_parser.py:1-67), on plain
argparse + PyYAML instead of configargparse so that the reference's `configs/*.yaml` load unchanged.

Precedence (configargparse semantics): command line > config file > argparse default.  Unknown keys in the
file are tolerated (the reference uses parse_known_args, config_parser.py:65; e.g. `nn_input_channel` in
default_config.yaml).  As in the reference, when `--config X` is given the default file is NOT read.
"""
import argparse
import os
import sys

import yaml

_SPEC = [
    # name, type, default, help      (order and defaults: config_parser.py:11-63)
    ("name", str, None, "name of the run, will be used to save/load checkpoints"),
    ("num_conv_layers", int, 8, "number of convolution layers"),
    ("ds_factor", int, 2, "downsample factor"),
    ("K", int, 2, "order of chebyshev polynomial for the VAE"),
    ("Kd", int, 3, "order of chebyshev polynomial for discriminator"),
    ("nf", int, 64, "number of conv filters of first encoder layer"),
    ("nz", int, 18, "Size of latent variable in latent space"),
    ("nz_cond", int, 24, "size of embedding of the first condition (pose)"),
    ("nz_cond2", int, 8, "size of embedding of the second condition (clotype)"),
    ("n_layer_cond", int, 1, "number of layers for the clothing type condition network"),
    ("activation", str, "b1leakyrelu", "b1relu, b1leakyrelu or b1tanh"),
    ("use_res_block", int, 0, "whether to use residual block in encoder"),
    ("use_res_block_dec", int, 1, "whether to use residual block in decoder"),
    ("cond_encoder", int, 0, "1 for condition the encoder too, 0 for not"),
    ("reduce_dim", int, 64, "reduce the channels in encoder final conv layer to this number"),
    ("affine", int, 0, "whether or not use affine residual block in decoder"),
    ("pose_type", str, "rot", "SMPL pose params or their rotational matrices"),
    ("optim_condnet", int, 1, ""),
    ("batch_size", int, 16, "input batch size for training"),
    ("num_epochs", int, 60, "number of training epochs"),
    ("lr", float, 8e-3, "Learning Rate"),
    ("lr_scaler", float, 1e-1, "lr is for G, multiply this scaler for D"),
    ("decay_every", int, 1, "decay lr after x epochs"),
    ("lr_warmup", int, 0, "Whether to use lr warmup"),
    ("seed", int, 123, "random seed"),
    ("restart", int, 1, "restart training or resume from checkpoint"),
    ("optimizer", str, "sgd", "adam or sgd (with momentum)"),
    ("loss", str, "l1", "l1 or l2"),
    ("loss_mask", str, "", "binary or None"),
    ("dataset", str, "dataset_male_4clotypes", "name of the dataset"),
    ("regularization", float, 2e-3, "weight for regularization term"),
    ("lambda_recon", float, 1.0, "coefficient for l1 loss"),
    ("lambda_edge", float, 1.0, "coefficient for edge loss"),
    ("lambda_latent", float, 8e-4, "coefficient for latent loss"),
    ("lambda_gan", float, 0.1, "coefficient for gan loss"),
    ("mode", str, "train", "train or test or demo"),
    ("gender", str, "male", "used to load smplx model of desired gender at test/demo"),
    ("smpl_model_folder", str, "body_models", "parent folder of the smpl model .pkl files"),
    ("demo_n_sample", int, 5, "generate n samples for demo"),
    ("save_obj", int, 1, "1 for saving meshes generated at demos"),
    ("vis_demo", int, 0, "1 for on-screen visualization"),
]
_CHOICES = {"pose_type": ["pose", "rot"], "optimizer": ["sgd", "adam"], "mode": ["train", "test", "demo"],
            "gender": ["female", "male"], "save_obj": [0, 1], "vis_demo": [0, 1]}
DEFAULT_CONFIG = "configs/default_config.yaml"


def _read_config_file(path):
    """Flat `key: value` file (configargparse's DefaultConfigFileParser accepts this YAML subset)."""
    with open(path) as f:
        data = yaml.safe_load(f) or {}
    if not isinstance(data, dict):
        raise ValueError("config file %s is not a flat key: value mapping" % path)
    return data


def parse_config(argv=None):
    """Returns (args, args_dict) like the reference (config_parser.py:65-68)."""
    argv = list(sys.argv[1:] if argv is None else argv)
    parser = argparse.ArgumentParser(prog="CAPE", description="CAPE model: mesh CVAE + discriminator",
                                     formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument("--config", default=DEFAULT_CONFIG, help="config file path")
    for name, typ, default, hlp in _SPEC:
        parser.add_argument("--" + name, type=typ, default=None, choices=_CHOICES.get(name), help=hlp)
    cli, _ = parser.parse_known_args(argv)
    values = {name: default for name, _, default, _ in _SPEC}
    cfg_path = cli.config
    file_vals = {}
    if cfg_path and os.path.exists(cfg_path):
        file_vals = _read_config_file(cfg_path)
    elif cfg_path and cfg_path != DEFAULT_CONFIG:
        raise FileNotFoundError(cfg_path)
    types = {name: typ for name, typ, _, _ in _SPEC}
    for k, v in file_vals.items():
        if k in values and v is not None:              # unknown keys tolerated (parse_known_args)
            v = types[k](v)
            if k in _CHOICES and v not in _CHOICES[k]:
                raise ValueError("invalid value %r for %s in %s" % (v, k, cfg_path))
            values[k] = v
    for name in values:
        v = getattr(cli, name)
        if v is not None:
            values[name] = v
    values["config"] = cfg_path
    args = argparse.Namespace(**values)
    return args, vars(args)


def model_params(args, n_train=None):
    """The kwargs main.py builds for models.CAPE (main.py:50-84), minus data-dependent entries."""
    p = dict(vars(args))
    p["restart"] = bool(args.restart)
    p["use_res_block"], p["use_res_block_dec"] = bool(args.use_res_block), bool(args.use_res_block_dec)
    p["nn_input_channel"] = 3
    nf = args.nf
    if args.num_conv_layers == 4:
        p["F"] = [nf, 2 * nf, 2 * nf, nf]
    elif args.num_conv_layers == 6:
        p["F"] = [nf, nf, 2 * nf, 2 * nf, 4 * nf, 4 * nf]
    elif args.num_conv_layers == 8:
        p["F"] = [nf, nf, 2 * nf, 2 * nf, 4 * nf, 4 * nf, 8 * nf, 8 * nf]
    else:
        raise NotImplementedError
    p["K"] = [2] * args.num_conv_layers           # main.py:65: --K is parsed but overridden
    p["Kd"] = args.Kd
    p["decay_steps"] = (args.decay_every * n_train / p["batch_size"]) if (args.mode == "train" and n_train) else 1
    p["cond_dim"], p["cond2_dim"] = 14 * 9, 4
    for k in ("cond_encoder", "affine", "lr_warmup", "optim_condnet"):
        p[k] = bool(getattr(args, k))
    for k in ("demo_n_sample", "mode", "dataset", "num_conv_layers", "ds_factor", "nf", "config", "pose_type",
              "decay_every", "gender", "save_obj", "vis_demo", "smpl_model_folder"):
        p.pop(k, None)
    return p
