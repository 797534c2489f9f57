"""Controllable-state embedding f(x) -> R^32 and its inverse-dynamics head.

Producer of the hot path's input (reference: ngu/models/intrinsic_novelty/
episodic_novelty/embedding.py:12-67; NGU paper appendix H.1).  Plain PyTorch on
cuDNN/cuBLAS: dense conv stack, out of the HBM-bound scan's scope (SURVEY.md
section 2, component 3).  Same architecture and optimiser hyper-parameters as
the reference so a trained state_dict transfers (module names ``siamese`` / ``h``).
"""
import torch
import torch.nn as nn
import torch.optim as optim

from .running_mean_std import RunningMeanStd


def _conv_out(size, kernel, stride):
    return (size - kernel) // stride + 1


def _orthogonal(layer, gain=nn.init.calculate_gain("relu")):
    nn.init.orthogonal_(layer.weight, gain)
    nn.init.zeros_(layer.bias)
    return layer


def atari_trunk(in_channels, out_features, out_gain=nn.init.calculate_gain("relu"), obs_hw=(84, 84)):
    """conv 8x8/4 -> 4x4/2 -> 3x3/1 -> linear, the trunk shared by Embedding and RND."""
    h, w = obs_hw
    for kern, stride in ((8, 4), (4, 2), (3, 1)):
        h, w = _conv_out(h, kern, stride), _conv_out(w, kern, stride)
    return nn.Sequential(
        _orthogonal(nn.Conv2d(in_channels, 32, 8, 4)), nn.ReLU(),
        _orthogonal(nn.Conv2d(32, 64, 4, 2)), nn.ReLU(),
        _orthogonal(nn.Conv2d(64, 64, 3, 1)), nn.ReLU(), nn.Flatten(),
        _orthogonal(nn.Linear(h * w * 64, out_features), out_gain))


class Embedding(nn.Module):
    def __init__(self, n_act, obs_shape, ctrl_state_dim, model_hypr, logger):
        super().__init__()
        self.n_act, self.obs_shape, self.ctrl_state_dim = n_act, obs_shape, ctrl_state_dim
        self.model_hypr, self.logger = model_hypr, logger
        self.ap_rms = RunningMeanStd(use_mpi=False)
        self.siamese = atari_trunk(obs_shape[0], ctrl_state_dim, obs_hw=obs_shape[1:])
        self.h = nn.Sequential(nn.ReLU(), _orthogonal(nn.Linear(2 * ctrl_state_dim, 128)), nn.ReLU(),
                               _orthogonal(nn.Linear(128, n_act), 0.01))
        self.optimizer = optim.Adam(self.parameters(),
                                    lr=model_hypr.get("learning_rate_action_prediction", 5e-4),
                                    betas=(model_hypr.get("adam_beta1", 0.9), model_hypr.get("adam_beta2", 0.999)),
                                    eps=model_hypr.get("adam_epsilon", 1e-4),
                                    weight_decay=model_hypr.get("action_prediction_l2_weight", 1e-5))
        self.criterion = nn.CrossEntropyLoss()
        self.update_count = 0
        self.grad_hook = None      # learner all-reduce seam (ngu_b200/learner_allreduce.py)

    def forward(self, obs_curr):
        return self.siamese(obs_curr)

    def step(self, timestep_obs, timestep_next_obs, timestep_act):
        """Maximum-likelihood inverse-dynamics update, one Adam step per timestep (embedding.py:51-67)."""
        params = list(self.parameters())
        for obs, nxt, act in zip(timestep_obs, timestep_next_obs, timestep_act):
            logits = self.h(torch.cat((self(obs), self(nxt)), dim=1))
            loss = self.criterion(logits, act)
            if self.grad_hook is not None:
                self.grad_hook.zero_grad()          # keeps .grad a view of the communication bucket
            else:
                self.optimizer.zero_grad()
            loss.backward()
            if self.grad_hook is not None:          # between backward and clip
                self.grad_hook(params)
            nn.utils.clip_grad_norm_(params, self.model_hypr.get("adam_clip_norm", 40))
            self.optimizer.step()
            self.ap_rms.update(loss.item())
        self.update_count += 1
        self.logger.log_scalar("ActionpredictionLoss", self.ap_rms.mean, self.update_count)
