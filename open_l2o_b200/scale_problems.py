"""Optimizee problems of the L2O-Scale drivers that the HierarchicalRNN benchmark needs (BASELINE config #4):
``ConvNet`` softmax classifier on image batches (SC/problems/problem_generator.py:429-479,637-696; SC/ =
Model_Free_L2O/L2O-Scale/L2O-Scale-Training/).  Forward/backward run through torch autograd on the device — the
optimizee is outside the learned-optimizer hot path."""
from __future__ import annotations

import torch
import torch.nn.functional as F


class ConvNet(object):
    """N-layer convnet for image classification (problem_generator.py:637-696): SAME-padded stride-1 convolutions
    with bias + activation, flatten, one dense layer — and, as in the reference, the activation is applied to the
    logits as well (:696).  ``image_shape`` = (channels, px, py); ``filter_list`` = [(kh, kw, n_out), ...].
    Parameters keep the reference's TF shapes: conv filters (kh, kw, c_in, c_out), dense (c*px*py, n_classes)."""

    def __init__(self, image_shape, n_classes, filter_list, activation=torch.relu, random_seed=None, noise_stdev=0.0):
        n_channels, px, py = image_shape
        self.activation = activation
        self.param_shapes = []
        c = n_channels
        for kh, kw, n_out in filter_list:
            self.param_shapes.append((kh, kw, c, n_out))
            self.param_shapes.append((n_out,))
            c = n_out
        self.affine_size = c * px * py
        self.param_shapes.append((self.affine_size, n_classes))
        self.param_shapes.append((n_classes,))
        self.random_seed = random_seed
        self.noise_stdev = noise_stdev

    def init_tensors(self, seed=None, device="cuda"):
        """tf.random_normal(shape, 0, 0.01) per parameter (:672-675)."""
        g = torch.Generator()
        if seed is not None:
            g.manual_seed(int(seed))
        return [(torch.randn(s, generator=g) * 0.01).to(device).requires_grad_(True) for s in self.param_shapes]

    def inference(self, params, data):
        """data: [batch, px, py, channels] (NHWC, as the reference feeds it)."""
        x = data.permute(0, 3, 1, 2)
        for i in range(0, len(params) - 2, 2):
            w, b = params[i], params[i + 1]
            kh, kw = w.shape[0], w.shape[1]
            x = F.conv2d(x, w.permute(3, 2, 0, 1), b, stride=1, padding=(kh // 2, kw // 2))   # SAME, odd kernels
            x = self.activation(x)
        flat = x.permute(0, 2, 3, 1).reshape(x.shape[0], self.affine_size)                 # NHWC flatten order
        return self.activation(flat @ params[-2] + params[-1])

    def objective(self, params, data, labels):
        """Softmax cross entropy averaged over the batch (:449-478); labels one-hot [batch, n_classes]."""
        logits = self.inference(params, data)
        return -(labels * F.log_softmax(logits, dim=1)).sum(1).mean()
