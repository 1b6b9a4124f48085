"""Training entry point with the reference's flags and config.ini semantics
(/root/reference/src/run.py:21-33, src/e2eflow/util.py:37-62, src/e2eflow/core/train.py:116-145):

    python -m unflow_b200.run --ex NAME [--debug] [--ow] [--config PATH] [--synthetic]
    python -m torch.distributed.run --nproc-per-node N -m unflow_b200.run --ex NAME   (data parallel)

* ``config.ini`` sections [dirs] [run] [train] [train_<dataset>] are parsed with the same type
  coercion (int -> float -> bool -> str); the dataset section overrides [train]
  (run.py:91-92); ``manual_decay_*`` strings become lists and define ``num_iters``.
* [run] batch_size is the TOTAL batch, divided by the number of GPUs (run.py:48).
* training runs in chunks of ``save_interval`` iterations; after each chunk a checkpoint
  ``model.ckpt-<iter>.pt`` (variables under their TF names, TF layout) is written and training
  resumes from the newest checkpoint by parsing the iteration from the file name
  (train.py:124-135, 258-259).  ``--ow`` discards an existing experiment, ``--debug`` disables
  checkpoint writing.
* The dataset adapters / TF queue-runner input pipeline of the reference are outside the hot path
  (SURVEY.md section 2): with ``--synthetic`` (or when [dirs] data does not exist) batches are
  seeded synthetic pairs of the configured height x width.
"""
import argparse
import configparser
import copy
import glob
import os
import shutil

import torch
import torch.distributed as dist

KITTI_NORMALIZATION = ([104.920005, 110.1753, 114.785955], 1 / 0.0039216)  # core/input.py:45-46


def config_dict(config_path):
    """util.py:37-62: the config as a dict of sections with intuitively typed values."""
    config = configparser.ConfigParser()
    if not config.read(config_path):
        raise FileNotFoundError(config_path)
    d = dict()
    for section_key in config.sections():
        sd = dict()
        section = config[section_key]
        for key in section:
            val = section[key]
            try:
                sd[key] = int(val)
            except ValueError:
                try:
                    sd[key] = float(val)
                except ValueError:
                    try:
                        sd[key] = section.getboolean(key)
                    except ValueError:
                        sd[key] = val
        d[section_key] = sd
    return d


def convert_input_strings(config_dct):
    """util.py:65-73 (the ``finetune`` checkpoint lookup is handled by load_finetune)."""
    if 'manual_decay_iters' in config_dct and 'manual_decay_lrs' in config_dct:
        iters_lst = [int(i) for i in str(config_dct['manual_decay_iters']).split(',')]
        lrs_lst = [float(l) for l in str(config_dct['manual_decay_lrs']).split(',')]
        config_dct['manual_decay_iters'] = iters_lst
        config_dct['manual_decay_lrs'] = lrs_lst
        config_dct['num_iters'] = sum(iters_lst)


def latest_checkpoint(ckpt_dir):
    best = None
    for p in glob.glob(os.path.join(ckpt_dir, "model.ckpt-*.pt")):
        try:
            it = int(os.path.basename(p)[len("model.ckpt-"):-3])
        except ValueError:
            continue
        if best is None or it > best[0]:
            best = (it, p)
    return best


def synthetic_batch(batch, height, width, step, rank, device):
    from . import synthetic
    im1, im2, _ = synthetic.image_pair(batch, height, width, seed=1234 + 7919 * step + rank)
    return im1.to(device), im2.to(device)


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument('--ex', default='default', help='Name of the experiment.')
    ap.add_argument('--debug', action='store_true', help='disable checkpoint writing for debugging')
    ap.add_argument('--ow', action='store_true', help='overwrite a previous experiment of the same name')
    ap.add_argument('--config', default=os.environ.get('UNFLOW_CONFIG', '../config.ini'))
    ap.add_argument('--synthetic', action='store_true')
    ap.add_argument('--max-iters', type=int, default=None, help='stop early (for smoke runs)')
    args = ap.parse_args(argv)

    cfg = config_dict(args.config)
    dirs, run_config = cfg.get('dirs', {}), cfg['run']
    train_dataset = run_config.get('dataset', 'kitti')
    params = copy.deepcopy(cfg['train'])
    params.update(cfg.get('train_' + train_dataset, {}))
    convert_input_strings(params)

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit("unflow_b200.run needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=device)
    assert run_config['batch_size'] % world == 0, 'batch_size must be divisible by the number of GPUs'
    gpu_batch_size = int(run_config['batch_size'] / world)

    log_dir = dirs.get('log', '../log')
    ckpt_dir = os.path.join(dirs.get('checkpoints', os.path.join(log_dir, 'checkpoints')), args.ex)
    if rank == 0:
        if args.ow and os.path.isdir(ckpt_dir):
            shutil.rmtree(ckpt_dir)
        os.makedirs(ckpt_dir, exist_ok=True)
    if world > 1:
        dist.barrier()

    from .e2eflow.core.train import Trainer
    tr = Trainer(params, KITTI_NORMALIZATION, device, seed=1234)

    num_iters = params.get('num_iters', 0)
    if args.max_iters is not None:
        num_iters = min(num_iters, args.max_iters)
    save_interval = min(params['save_interval'], max(num_iters, 1))
    start_iter = 1
    ck = latest_checkpoint(ckpt_dir)
    if ck is not None:
        state = torch.load(ck[1], map_location='cpu')
        tr.variables.load_tf_dict(state['variables'])
        tr.adam_m.copy_(state['adam_m']); tr.adam_v.copy_(state['adam_v'])
        tr.iteration = ck[0]
        start_iter = ck[0] + 1
        if start_iter > num_iters:
            print('-- train: max_iter reached')
            return
    tr.broadcast_variables(0)
    if rank == 0:
        print('-- training from i = {} to {}'.format(start_iter, num_iters))

    data_dir = dirs.get('data', '')
    if not args.synthetic and os.path.isdir(data_dir):
        raise SystemExit("real-data input pipelines are outside the hot path; run with --synthetic")
    for i in range(start_iter, num_iters + 1):
        im1, im2 = synthetic_batch(gpu_batch_size, params['height'], params['width'], i, rank, device)
        loss = tr.step(im1, im2)     # LR schedule inside (train.py:225-244)
        if rank == 0 and (i == 1 or i % params['display_interval'] == 0):
            print("-- train: i = {}, loss = {}".format(i, float(loss)))
        if i % save_interval == 0 and not args.debug and rank == 0:
            torch.save({'variables': tr.variables.to_tf_dict(), 'adam_m': tr.adam_m.cpu(),
                        'adam_v': tr.adam_v.cpu()}, os.path.join(ckpt_dir, 'model.ckpt-%d.pt' % i))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
