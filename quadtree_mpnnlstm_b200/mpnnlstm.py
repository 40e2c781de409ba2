"""Trainer around the seq2seq model: drop-in for ``model/mpnnlstm.py`` of the reference (``NextFramePredictorS2S``,
model/mpnnlstm.py:81-443) -- same constructor, ``train`` / ``predict`` / ``save`` / ``load`` / ``get_climatology_array``
signatures, same loss (MSE, or BCE for ``binary``, on the unmasked pixels), ``clip_grad_norm_(10)``, Adam + StepLR(3),
TensorBoard tags and ``.pth`` format -- so ``ice_exp.py`` / ``ice_inf.py`` run against it unchanged.

What is different is where the work happens:
  * the per-sample host work of the reference loop is gone: the unmasked-pixel index lives on the device and is built
    once (the reference indexes a device tensor with a NumPy mask every step), losses are accumulated on the device and
    read back once per epoch (the reference calls ``.item()`` and ``torch.cuda.empty_cache()`` per sample), climatology
    rows are gathered on the device;
  * on the pixel-wise mesh (``thresh = -inf``, the ice_exp.py default) and with ``use_cuda_graph=True`` the whole
    optimizer step -- forward, loss on the mesh nodes (one pixel per node: the same numbers), backward, clip, Adam --
    is one CUDA-graph replay (``train.TrainStep``), its learning rate a device scalar that follows the StepLR schedule;
  * ``DeviceWindowDataset`` keeps ONE [T, H, W, c] cube on the device and serves (x, y, launch_date) sliding windows as
    views, instead of materialising every window on the host (ice_dataset.py:20-68).

Truncated back-propagation (``truncated_backprop > 0``, the reference's default of 45; model/mpnnlstm.py:281-313) runs the
reference's chunked loop as it is written there, quirks included: every chunk re-encodes the inputs and unrolls
``range(t - truncated_backprop, t)`` from the encoder state, ``zero_grad`` runs at the start of EVERY chunk (so the optimizer
step after the last chunk sees the last chunk's gradients only), no gradient clipping, and a chunk that runs past
``output_timesteps`` raises ``IndexError`` (``output_timesteps`` must be a multiple of ``truncated_backprop``).
"""
from __future__ import annotations

import datetime
import os
import time
from abc import ABC, abstractmethod

import numpy as np
import torch

from .graph_functions import image_to_graph, unflatten
from .seq2seq import Seq2Seq
from .utils import add_positional_encoding, get_n_params, int_to_datetime


class _NullWriter:
    """Stands in for torch.utils.tensorboard.SummaryWriter when tensorboard is not installed."""

    def add_scalar(self, *a, **k):
        pass

    def flush(self):
        pass


def _summary_writer(path):
    try:
        from torch.utils.tensorboard import SummaryWriter
        return SummaryWriter(path)
    except Exception:       # tensorboard missing or unwritable run directory: training must not depend on it
        return _NullWriter()


class DeviceWindowDataset:
    """Sliding (x, y, launch_date) windows over one device-resident cube.

    cube [T, H, W, c] (float32, already normalised), ``y_channels`` = channels of the target, ``times`` = int64
    nanosecond timestamps of the T frames (``launch_date`` of window i is ``times[i + input_timesteps]``, as in
    ice_dataset.py:60).  Iterating yields the reference loader's batch-of-one layout: x [1, T_in, H, W, c],
    y [1, T_out, H, W, len(y_channels)], launch_date int64 tensor [1] (on the host, like a DataLoader's)."""

    def __init__(self, cube, input_timesteps, output_timesteps, times=None, y_channels=(0,), indices=None, shuffle=False,
                 seed=0):
        assert cube.dim() == 4, "cube must be [T, H, W, c]"
        self.cube = cube
        self.t_in, self.t_out = int(input_timesteps), int(output_timesteps)
        self.y_channels = list(y_channels)
        n = cube.shape[0] - self.t_in - self.t_out + 1
        self.indices = list(range(max(n, 0))) if indices is None else [int(i) for i in indices]
        self.times = (torch.arange(cube.shape[0], dtype=torch.int64) * 86_400_000_000_000 if times is None
                      else torch.as_tensor(np.asarray(times).astype("int64")))
        self.image_shape = tuple(cube.shape[1:3])
        self.shuffle, self._rng = shuffle, np.random.default_rng(seed)
        self.dataset = self                     # the trainer reads loader.dataset.image_shape

    def __len__(self):
        return len(self.indices)

    def window(self, i):
        a = self.indices[i]
        x = self.cube[a:a + self.t_in]
        y = self.cube[a + self.t_in:a + self.t_in + self.t_out][..., self.y_channels]
        return x.unsqueeze(0), y.unsqueeze(0), self.times[a + self.t_in].reshape(1)

    def __iter__(self):
        order = self._rng.permutation(len(self.indices)) if self.shuffle else range(len(self.indices))
        for i in order:
            yield self.window(int(i))


class NextFramePredictor(ABC):
    def __init__(self, thresh, experiment_name='experiment', decompose=True, input_features=1, transform_func=None,
                 condition='max_larger_than', device=None):
        self.experiment_name = experiment_name
        self.decompose = decompose
        self.model = None
        self.thresh = thresh
        self.transform_func = transform_func
        self.condition = condition
        self.input_features = input_features
        self.device = device

    @abstractmethod
    def train(self, *a, **k):
        pass

    @abstractmethod
    def predict(self, *a, **k):
        pass

    def score(self, x, y, rollout=None):
        pass


class NextFramePredictorS2S(NextFramePredictor):
    def __init__(self, thresh, experiment_name='experiment', decompose=True, input_features=1, input_timesteps=3,
                 output_timesteps=3, device=None, transform_func=None, condition='max_larger_than', remesh_input=False,
                 binary=False, debug=False, model_kwargs={}, use_cuda_graph=False):
        super().__init__(thresh=thresh, experiment_name=experiment_name, decompose=decompose, input_features=input_features,
                         device=device, transform_func=transform_func, condition=condition)
        self.input_timesteps = input_timesteps
        self.output_timesteps = output_timesteps
        self.binary = binary
        self.thresh = thresh if decompose else -np.inf
        self.debug = debug
        self.use_cuda_graph = use_cuda_graph
        # + 3: positional encoding (x, y) and node size (model/mpnnlstm.py:124)
        self.model = Seq2Seq(input_features=input_features + 3, input_timesteps=input_timesteps,
                             output_timesteps=output_timesteps, thresh=thresh, device=device, remesh_input=remesh_input,
                             binary=binary, debug=debug, **model_kwargs).to(device)
        self.training_initiated = False
        self._keep_cache = None
        self._graph_step = None

    # ---- bookkeeping (model/mpnnlstm.py:158-185) ---------------------------------------------------------------
    def get_n_params(self):
        return get_n_params(self.model)

    def save(self, directory):
        torch.save(self.model.state_dict(), os.path.join(directory, f'{self.experiment_name}.pth'))

    def load(self, directory):
        path = os.path.join(directory, f'{self.experiment_name}.pth')
        try:
            self.model.load_state_dict(torch.load(path))
        except Exception:
            self.model.load_state_dict(torch.load(path, map_location=torch.device('cpu')))

    def initiate_training(self, lr, lr_decay):
        from torch.optim.lr_scheduler import StepLR
        self.loss_func = torch.nn.MSELoss() if not self.binary else torch.nn.BCELoss()
        self.loss_func_name = 'MSE' if not self.binary else 'BCE'
        self.optimizer = torch.optim.Adam(self.model.parameters(), lr=lr)
        self.scheduler = StepLR(self.optimizer, step_size=3, gamma=lr_decay)
        self.writer = _summary_writer('runs/' + self.experiment_name + '_' + datetime.datetime.now().strftime("%Y%m%d_%H_%M_%S"))
        self.test_loss = []
        self.train_loss = []
        self.training_initiated = True

    # ---- helpers -----------------------------------------------------------------------------------------------
    def _keep(self, mask, image_shape, device):
        """Device index of the unmasked pixels (flattened H*W), built once per mask."""
        key = (id(mask), tuple(image_shape), str(device))
        if self._keep_cache is None or self._keep_cache[0] != key:
            m = np.zeros(image_shape, bool) if mask is None else np.asarray(mask.cpu() if torch.is_tensor(mask) else mask, dtype=bool)
            self._keep_cache = (key, torch.from_numpy(np.flatnonzero(~m.reshape(-1))).to(device))
        return self._keep_cache[1]

    def _loss(self, y_hat, mappings, y, image_shape, mask):
        """loss_func(y_hat[:, ~mask], y[:, ~mask]) of the reference (model/mpnnlstm.py:243-246)."""
        keep = self._keep(mask, image_shape, y.device)
        T = len(y_hat)
        img = torch.stack([unflatten(y_hat[i], mappings[i], image_shape, mask) for i in range(T)], dim=0)
        a = img.reshape(T, -1, img.shape[-1]).index_select(1, keep)
        b = y.reshape(T, -1, y.shape[-1]).index_select(1, keep)
        return self.loss_func(a, b.to(a.dtype))

    def get_climatology_array(self, climatology, launch_date):
        """Daily climate normals of the output timesteps: climatology [n_vars, 365|366, H, W] -> [T_out, H, W, n_vars]
        (model/mpnnlstm.py:389-400), gathered on the device the climatology lives on."""
        t0 = int(np.asarray(launch_date.cpu() if torch.is_tensor(launch_date) else launch_date).reshape(-1)[0])
        doys = [int_to_datetime(t0 + 8.640e13 * t).timetuple().tm_yday - 1 for t in range(0, self.output_timesteps)]
        if not torch.is_tensor(climatology):
            climatology = torch.as_tensor(np.asarray(climatology))
        out = climatology.index_select(1, torch.as_tensor(doys, device=climatology.device))
        return torch.moveaxis(out, 0, -1)

    # ---- training (model/mpnnlstm.py:187-387) ------------------------------------------------------------------
    def train(self, loader_train, loader_test, climatology=None, n_epochs=200, lr=0.01, lr_decay=0.95, mask=None,
              high_interest_region=None, truncated_backprop=45, graph_structure=None):
        import pandas as pd
        image_shape = tuple(loader_train.dataset.image_shape)
        if not self.training_initiated:
            self.initiate_training(lr, lr_decay)
        if mask is not None:
            assert tuple(mask.shape) == image_shape, f'Mask and image shapes do not match. Got {mask.shape} and {image_shape}'
        dev = self.device
        params = list(self.model.parameters())
        st = time.time()
        batch_step = 0
        for epoch in range(n_epochs):
            running = torch.zeros((), device=dev)
            step = 0
            for x, y, launch_date in loader_train:
                x, y = x.squeeze(0).to(dev), y.squeeze(0).to(dev)
                concat_layers = self.get_climatology_array(climatology, launch_date) if climatology is not None else None
                if truncated_backprop == 0:
                    loss = self._train_step(x, y, concat_layers, mask, high_interest_region, graph_structure, image_shape, params,
                                            batch_step)
                else:
                    loss = self._train_step_truncated(x, y, concat_layers, mask, high_interest_region, graph_structure,
                                                      image_shape, truncated_backprop)
                self.writer.add_scalar("Loss/train", loss, batch_step)      # unconditional, per batch (model/mpnnlstm.py:315)
                running = running + loss
                step += 1
                batch_step += 1
            running_test = torch.zeros((), device=dev)
            step_test = 0
            for x, y, launch_date in loader_test:
                x, y = x.squeeze(0).to(dev), y.squeeze(0).to(dev)
                concat_layers = self.get_climatology_array(climatology, launch_date) if climatology is not None else None
                with torch.no_grad():
                    y_hat, maps = self.model(x, y, concat_layers, teacher_forcing_ratio=0, mask=mask,
                                             high_interest_region=high_interest_region, graph_structure=graph_structure)
                    running_test = running_test + self._loss(y_hat, maps, y, image_shape, mask)
                step_test += 1
            # one read-back per epoch (the reference's own (step + 1) denominators, model/mpnnlstm.py:359-360)
            running_loss, running_loss_test = (running / (step + 1)).item(), (running_test / (step_test + 1)).item()
            if np.isnan(running_loss_test):
                raise ValueError('NaN loss :(')
            if running_loss_test > 4:
                raise ValueError('Diverged :(')
            self.writer.add_scalar("Loss/train_epoch", running_loss, epoch)
            self.writer.add_scalar("Loss/test", running_loss_test, epoch)
            self.scheduler.step()
            self.train_loss.append(running_loss)
            self.test_loss.append(running_loss_test)
            print(f"{self.experiment_name} | Epoch {epoch} train {self.loss_func_name}: {running_loss:.4f}, "
                  f"test {self.loss_func_name}: {running_loss_test:.4f}, lr: {self.scheduler.get_last_lr()[0]:.4f}, "
                  f"time_per_epoch: {(time.time() - st) / (epoch + 1):.1f}")
        print(f'Finished in {(time.time() - st) / 60} minutes')
        self.writer.flush()
        self.loss = pd.DataFrame({'train_loss': self.train_loss, 'test_loss': self.test_loss})

    def _train_step_truncated(self, x, y, concat_layers, mask, hir, graph_structure, image_shape, truncated_backprop):
        """The reference's chunked loop (model/mpnnlstm.py:281-313), statement for statement."""
        output_timestep = 0
        loss = None
        while output_timestep < self.output_timesteps:
            output_timestep = min(output_timestep + truncated_backprop, self.output_timesteps + 1)
            unroll_steps = range(output_timestep - truncated_backprop, output_timestep)
            self.optimizer.zero_grad()
            self.model.process_inputs(x, mask=mask, high_interest_region=hir, graph_structure=graph_structure)
            y_hat, maps = self.model.unroll_output(unroll_steps, y, concat_layers=concat_layers, teacher_forcing_ratio=0, mask=mask,
                                                   high_interest_region=hir, remesh_every=1)
            loss = self._loss(y_hat, maps, y[list(unroll_steps)], image_shape, mask)
            loss.backward(retain_graph=True)
            del y_hat, maps
        self.optimizer.step()
        return loss.detach()

    def _train_step(self, x, y, concat_layers, mask, hir, graph_structure, image_shape, params, batch_step=0):
        # node-space loss == pixel-space loss only on the pixel-wise mesh (one pixel per node), which is what TrainStep captures
        if (self.use_cuda_graph and self.model.thresh == -float("inf") and graph_structure is None and not self.binary
                and concat_layers is not None and x.is_cuda and hir is None):
            return self._graph_train_step(x, y, concat_layers, mask, graph_structure)
        self.optimizer.zero_grad()
        y_hat, maps = self.model(x, y, concat_layers, teacher_forcing_ratio=0, mask=mask, high_interest_region=hir,
                                 graph_structure=graph_structure)
        loss = self._loss(y_hat, maps, y, image_shape, mask)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, max_norm=10)
        self.optimizer.step()
        if self.debug:
            for name, mod in (("encoder", self.model.encoder), ("decoder", self.model.decoder)):
                g = [torch.norm(p.grad.detach()) for p in mod.parameters() if p.grad is not None]
                if g:
                    self.writer.add_scalar(f"Grad/{name}/grad_norms", torch.norm(torch.stack(g)), batch_step)
        return loss.detach()

    def _graph_train_step(self, x, y, concat_layers, mask, graph_structure):
        """Static mesh: the whole optimizer step (forward, node-space MSE, backward, clip, Adam) as one CUDA-graph replay;
        the learning rate lives in a device scalar that follows this trainer's scheduler."""
        from .train import TrainStep
        if self._graph_step is None:
            lr = self.optimizer.param_groups[0]["lr"]
            self._graph_step = TrainStep(self.model, mask, lr=torch.tensor(float(lr), device=x.device),
                                         graph_structure=graph_structure, use_cuda_graph=True)
        ts = self._graph_step
        ts.opt.param_groups[0]["lr"].fill_(float(self.optimizer.param_groups[0]["lr"]))
        return ts(x, y, concat_layers)

    # ---- inference (model/mpnnlstm.py:402-440) -----------------------------------------------------------------
    def predict(self, loader, climatology=None, mask=None, high_interest_region=None, graph_structure=None):
        image_shape = tuple(loader.dataset.image_shape)
        self.model.to(self.device)
        y_pred = []
        for x, y, launch_date in loader:
            x = x.squeeze(0).to(self.device)
            concat_layers = self.get_climatology_array(climatology, launch_date) if climatology is not None else None
            with torch.no_grad():
                y_hat, maps = self.model(x, concat_layers=concat_layers, teacher_forcing_ratio=0, mask=mask,
                                         high_interest_region=high_interest_region, graph_structure=graph_structure)
                y_pred.append(torch.stack([unflatten(y_hat[i], maps[i], image_shape, mask)
                                           for i in range(self.output_timesteps)]))
        return torch.stack(y_pred, 0).cpu().numpy()       # ONE device -> host copy for the whole loader

    def test_threshold(self, x, thresh, mask=None, high_interest_region=None, contours=True):
        """Mesh preview (model/mpnnlstm.py:138-156): returns (reconstructed images [n, H, W], labels, number of nodes); the
        reference draws them with matplotlib, which stays with the caller here."""
        n_sample, w, h, c = x.shape
        graph = image_to_graph(add_positional_encoding(x), thresh=thresh, mask=mask, high_interest_region=high_interest_region,
                               transform_func=self.transform_func)
        img = unflatten(graph['data'][..., [0]], graph['mapping'], (w, h))
        return img[..., 0], graph.get('labels'), int(graph['data'].shape[1])
