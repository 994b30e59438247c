"""Checkpoint files and the per-epoch learning-rate rule: the two pieces of training state that cross the boundary of
the fused iteration (scope row f-1's callers: ``train.py`` -> ``model.setup`` / ``save_networks`` / ``update_learning_rate``).

    save_networks / load_networks / latest_epoch   models/base_model.py:75-90, 181-219: one ``<epoch>_net_<Name>.pth`` per
                                                   net = ``torch.save(net.state_dict())`` of the *unwrapped* module (no
                                                   ``module.`` prefix), names Coarse / Fine (models/nerf_downX_model.py:176);
                                                   files are interchangeable with the reference in both directions
    lr_at_epoch / LrSchedule                       models/networks.py:88-118 (``get_scheduler``): the 'linear' / 'exp'
                                                   LambdaLR rules and StepLR, as the learning rate in force after ``epoch``
                                                   scheduler steps -- what ``Trainer.optimize_parameters(lr=...)`` takes

torch is used for (de)serialisation only: ``.pth`` is torch's own container format."""
from __future__ import annotations

import glob
import math
import os
import re
from collections import OrderedDict
from typing import Dict, Mapping, Optional, Tuple

import torch

from .renderer import state_dict_order

NET_NAMES = ("Coarse", "Fine")


def _net_path(save_dir: str, epoch, name: str) -> str:
    return os.path.join(save_dir, "%s_net_%s.pth" % (epoch, name))


def save_networks(save_dir: str, epoch, state_coarse: Mapping[str, torch.Tensor], state_fine: Mapping[str, torch.Tensor]) -> Tuple[str, str]:
    """Write ``<epoch>_net_Coarse.pth`` / ``<epoch>_net_Fine.pth`` (``epoch`` may be 'latest').  Tensors are saved on the
    CPU in state_dict registration order, like ``net.cpu().state_dict()``."""
    os.makedirs(save_dir, exist_ok=True)
    paths = []
    for name, sd in zip(NET_NAMES, (state_coarse, state_fine)):
        D = sum(1 for k in sd if re.fullmatch(r"xyz_encoding_\d+\.0\.weight", k))
        order = [k for k in state_dict_order(D) if k in sd] + [k for k in sd if k not in state_dict_order(D)]
        out = OrderedDict((k, sd[k].detach().to("cpu").contiguous().clone()) for k in order)
        path = _net_path(save_dir, epoch, name)
        torch.save(out, path)
        paths.append(path)
    return paths[0], paths[1]


def load_networks(save_dir: str, epoch, keys: Optional[str] = None, map_location="cpu") -> Tuple[Dict[str, torch.Tensor], Dict[str, torch.Tensor]]:
    """Read the two state dicts back.  ``keys``: the reference's ``--init_weights_keys`` regular expression (only matching
    entries are returned, models/base_model.py:216-218).  A ``module.`` prefix (a checkpoint written from a wrapped net by
    other tooling) is stripped."""
    out = []
    for name in NET_NAMES:
        path = _net_path(save_dir, epoch, name)
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        sd = torch.load(path, map_location=map_location, weights_only=True)
        sd = OrderedDict((k[len("module."):] if k.startswith("module.") else k, v) for k, v in sd.items())
        if keys is not None:
            sd = OrderedDict((k, v) for k, v in sd.items() if re.match(keys, k))
        out.append(sd)
    return out[0], out[1]


def latest_epoch(save_dir: str) -> int:
    """``--load_epoch latest`` (models/base_model.py:87-88): the largest numeric epoch among ``*.pth`` in the directory."""
    epochs = [int(os.path.basename(x).split("_")[0]) for x in glob.glob(os.path.join(save_dir, "*.pth")) if "latest" not in x]
    if not epochs:
        raise FileNotFoundError(f"no numbered checkpoints in {save_dir}")
    return max(epochs)


def lr_at_epoch(epoch: int, lr: float = 5e-4, lr_final: float = 5e-6, n_epochs: int = 20, n_epochs_decay: int = 10,
                lr_policy: str = "exp", lr_decay_epochs: int = 10, lr_decay_gamma: float = 0.1) -> float:
    """Learning rate after ``epoch`` scheduler steps (``epoch`` = 0 during the first epoch): models/networks.py:102-115.
    'linear' / 'exp' hold ``lr`` for ``n_epochs - n_epochs_decay`` epochs and then move to ``lr_final`` linearly / log-linearly
    over ``n_epochs_decay + 1`` steps; 'step' multiplies by ``lr_decay_gamma`` every ``lr_decay_epochs`` epochs."""
    if lr_policy in ("linear", "exp"):
        t = max(0, epoch + 1 - n_epochs + n_epochs_decay) / float(n_epochs_decay + 1)
        if lr_policy == "linear":
            cur = lr * (1 - t) + lr_final * t
        else:
            cur = math.exp(math.log(lr) * (1 - t) + math.log(lr_final) * t)
        return (cur / lr) * lr          # LambdaLR multiplies the initial rate by the rule's ratio
    if lr_policy == "step":
        return lr * lr_decay_gamma ** (epoch // lr_decay_epochs)
    raise NotImplementedError("learning rate policy [%s] is not implemented" % lr_policy)


class LrSchedule:
    """The scheduler object of ``model.setup`` + ``update_learning_rate`` (models/base_model.py:103-104, 154-159) for
    a ``Trainer``: ``lr`` is the rate to pass to ``optimize_parameters``; ``step()`` ends an epoch.

    ``current_epoch`` > 0 is ``--continue_train``.  The LambdaLR policies are closed-form in the epoch, so a resumed run
    continues the schedule.  torch's StepLR is a recurrence on the optimiser's *current* rate, and the reference builds a
    fresh optimiser at ``opt.lr`` on resume, so under 'step' a resumed run restarts from ``lr`` and decays at the next
    multiples of ``lr_decay_epochs`` -- mirrored here as the same recurrence."""

    def __init__(self, lr: float = 5e-4, lr_final: float = 5e-6, n_epochs: int = 20, n_epochs_decay: int = 10,
                 lr_policy: str = "exp", lr_decay_epochs: int = 10, lr_decay_gamma: float = 0.1, current_epoch: int = 0):
        self.kw = dict(lr=lr, lr_final=lr_final, n_epochs=n_epochs, n_epochs_decay=n_epochs_decay, lr_policy=lr_policy,
                       lr_decay_epochs=lr_decay_epochs, lr_decay_gamma=lr_decay_gamma)
        self.epoch = int(current_epoch)            # get_scheduler(last_epoch=current_epoch - 1) then one implicit step
        lr_at_epoch(self.epoch, **self.kw)         # validates the policy name
        self._step_lr = lr * (lr_decay_gamma if (lr_policy == "step" and self.epoch > 0 and self.epoch % lr_decay_epochs == 0) else 1.0)

    @classmethod
    def from_opt(cls, opt, current_epoch: int = 0) -> "LrSchedule":
        return cls(opt.lr, opt.lr_final, opt.n_epochs, opt.n_epochs_decay, opt.lr_policy, getattr(opt, "lr_decay_epochs", 10),
                   getattr(opt, "lr_decay_gamma", 0.1), current_epoch)

    @property
    def lr(self) -> float:
        if self.kw["lr_policy"] == "step":
            return self._step_lr
        return lr_at_epoch(self.epoch, **self.kw)

    def step(self) -> float:
        self.epoch += 1
        if self.kw["lr_policy"] == "step" and self.epoch % self.kw["lr_decay_epochs"] == 0:
            self._step_lr *= self.kw["lr_decay_gamma"]
        return self.lr
