"""RND life-long novelty: producer of the modulator alpha_t.

Reference: ngu/models/intrinsic_novelty/lifelong_novelty/{lifelong_novelty,rnd_prediction}.py.
The two CNN forwards are plain PyTorch/cuDNN (out of the scan's scope); the
scalar tail err -> running mean/std -> alpha = 1 + (err - mu)/sigma follows
lifelong_novelty.py:53-57.  The clip to [1, L] and the product with the
episodic reward are fused into the engine's finish kernel.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.optim as optim

from . import pytorch_util as ptu
from .embedding import atari_trunk
from .running_mean_std import RunningMeanStd


class RNDPrediction(nn.Module):
    """Feature network of RND (NGU paper appendix H.2): Atari trunk -> 128 features."""

    def __init__(self, obs_shape=(1, 84, 84), model_hypr=None):
        super().__init__()
        self.obs_shape, self.model_hypr = obs_shape, model_hypr
        trunk = atari_trunk(obs_shape[0], 128, out_gain=0.01, obs_hw=obs_shape[1:])
        # attribute names follow the reference so checkpoints load: conv1..3, fc
        self.conv1, self.conv2, self.conv3, self.flatten, self.fc = trunk[0], trunk[2], trunk[4], trunk[6], trunk[7]

    def forward(self, obs):
        x = torch.relu(self.conv1(obs))
        x = torch.relu(self.conv2(x))
        x = torch.relu(self.conv3(x))
        return self.fc(self.flatten(x))


class LifelongNovelty(nn.Module):
    def __init__(self, obs_shape, model_hypr, logger):
        super().__init__()
        self.obs_shape, self.model_hypr, self.logger = obs_shape, model_hypr, logger
        self._host_ll_rms = RunningMeanStd(use_mpi=False)
        self._device_ll_rms = None      # set by IntrinsicNovelty when the fused (on-device) tail is in use
        self.rnd_loss_rms = RunningMeanStd(use_mpi=False)
        self.predictor = RNDPrediction(obs_shape, model_hypr)
        self.target = RNDPrediction(obs_shape, model_hypr)
        for p in self.target.parameters():
            p.requires_grad = False
        self.optimizer = optim.Adam(self.predictor.parameters(), lr=model_hypr.get("learning_rate_rnd", 5e-4),
                                    betas=(model_hypr.get("adam_beta1", 0.9), model_hypr.get("adam_beta2", 0.999)),
                                    eps=model_hypr.get("adam_epsilon", 1e-4))
        self.criterion = nn.MSELoss()
        self.update_count = 0
        self.grad_hook = None

    @property
    def ll_rms(self):
        """Running mean/std of the RND error (lifelong_novelty.py:26).  Lives on the GPU when
        IntrinsicNovelty runs the fused tail (ngu_epi_step_rnd), on the host otherwise."""
        return self._device_ll_rms if self._device_ll_rms is not None else self._host_ll_rms

    def forward(self, obs):
        return self.predictor(obs), self.target(obs)

    @torch.no_grad()
    def compute_lifelong_curiosity(self, obs):
        """alpha_t (lifelong_novelty.py:49-57); returns a CPU float64 tensor like the reference."""
        pred, targ = self(obs.to(ptu.current_device()))
        err = (pred - targ).pow(2).mean(dim=1).cpu().numpy()
        rms = self._host_ll_rms
        rms.update(err)
        return torch.from_numpy(1.0 + (err - rms.mean) / np.sqrt(rms.var))

    def step(self, timestep_obs):
        params = list(self.predictor.parameters())
        for obs in timestep_obs:
            pred, targ = self(obs)
            loss = self.criterion(targ, pred)
            if self.grad_hook is not None:
                self.grad_hook.zero_grad()          # keeps .grad a view of the communication bucket
            else:
                self.optimizer.zero_grad()
            loss.backward()
            if self.grad_hook is not None:
                self.grad_hook(params)
            nn.utils.clip_grad_norm_(params, self.model_hypr.get("adam_clip_norm", 40))
            self.optimizer.step()
            self.rnd_loss_rms.update(loss.item())
        self.update_count += 1
        self.logger.log_scalar("RNDLoss", self.rnd_loss_rms.mean, self.update_count)
        self.logger.log_scalar("LifelongNoveltyMean", self.ll_rms.mean, self.update_count)
        self.logger.log_scalar("LifelongNoveltyVar", self.ll_rms.var, self.update_count)

    def to(self, device):
        self.predictor.to(device)
        self.target.to(device)
        return self
