"""Parameter containers for the dilated causal TCN (reference: scripts/model/tcn.py:7-64).

The classes keep the reference's constructor signatures, sub-module names and therefore state_dict keys (including the
duplicate `net.0.* / net.4.*` aliases of conv1 / conv2, tcn.py:30-31) so that reference checkpoints load with
strict=True.  They hold parameters only: the arithmetic (weight-norm, left-padded dilated conv without the chomp copy,
ReLU, dropout, residual) is executed by the sm_100a kernels in tgb200.engine.GeneratorEngine.text_forward; calling
forward() on these containers directly raises, there is no PyTorch fallback."""
import torch.nn as nn
from torch.nn.utils import weight_norm


class Chomp1d(nn.Module):
    def __init__(self, chomp_size):
        super().__init__()
        self.chomp_size = chomp_size

    def forward(self, x):
        raise RuntimeError('Chomp1d is fused away (causal left padding inside tg_conv_gemm_f32); run the parent PoseGenerator')


class TemporalBlock(nn.Module):
    def __init__(self, n_inputs, n_outputs, kernel_size, stride, dilation, padding, dropout=0.2):
        super().__init__()
        assert stride == 1
        self.kernel_size, self.dilation, self.padding = kernel_size, dilation, padding
        self.conv1 = weight_norm(nn.Conv1d(n_inputs, n_outputs, kernel_size, stride=stride, padding=padding, dilation=dilation))
        self.chomp1, self.relu1, self.dropout1 = Chomp1d(padding), nn.ReLU(), nn.Dropout(dropout)
        self.conv2 = weight_norm(nn.Conv1d(n_outputs, n_outputs, kernel_size, stride=stride, padding=padding, dilation=dilation))
        self.chomp2, self.relu2, self.dropout2 = Chomp1d(padding), nn.ReLU(), nn.Dropout(dropout)
        self.net = nn.Sequential(self.conv1, self.chomp1, self.relu1, self.dropout1, self.conv2, self.chomp2, self.relu2, self.dropout2)
        self.downsample = nn.Conv1d(n_inputs, n_outputs, 1) if n_inputs != n_outputs else None
        self.relu = nn.ReLU()
        self.init_weights()

    def init_weights(self):
        # same draws as tcn.py:37-41 (they touch .weight, which weight_norm recomputes from g, v: a no-op on the model)
        self.conv1.weight.data.normal_(0, 0.01)
        self.conv2.weight.data.normal_(0, 0.01)
        if self.downsample is not None:
            self.downsample.weight.data.normal_(0, 0.01)

    def forward(self, x):
        raise RuntimeError('TemporalBlock holds parameters only; its kernels run inside PoseGenerator.forward')


class TemporalConvNet(nn.Module):
    def __init__(self, num_inputs, num_channels, kernel_size=2, dropout=0.2):
        super().__init__()
        layers = []
        for i, out_channels in enumerate(num_channels):
            dilation = 2 ** i
            in_channels = num_inputs if i == 0 else num_channels[i - 1]
            layers.append(TemporalBlock(in_channels, out_channels, kernel_size, stride=1, dilation=dilation,
                                        padding=(kernel_size - 1) * dilation, dropout=dropout))
        self.network = nn.Sequential(*layers)

    def forward(self, x):
        raise RuntimeError('TemporalConvNet holds parameters only; its kernels run inside PoseGenerator.forward')
