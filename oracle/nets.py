"""TEST INFRASTRUCTURE ONLY (oracle) — CPU PyTorch restatement of PEPPER's two inference networks and their
predict loops.  Never imported by the product package.

  VariantNet  <-> pepper_variant/modules/python/models/simple_model.py:6-82  (class TransducerGRU, nn.LSTM based)
  PolishNet   <-> pepper/modules/python/models/simple_model.py:5-42          (class TransducerGRU, nn.GRU based)
  polish_predict  <-> pepper/modules/python/models/predict_distributed_cpu.py:50-90 (the 19-window loop; the
                      reference runs the same module through onnxruntime there, predict.py:47-93 is the torch loop)
  variant_predict <-> pepper_variant/modules/python/models/predict_distributed_cpu.py:102-147 (predict_pytorch)

The restatement keeps the reference's parameter names, so a reference ``state_dict`` loads unchanged.  It is pinned
against the reference classes imported from /root/reference by tests/golden/make_golden_nets.py (run in the build
container; torch version recorded in the fixture) and tests/test_oracle_nets.py.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn


class VariantNet(nn.Module):
    def __init__(self, image_features=26, hidden=256, window=33, num_classes_type=3):
        super().__init__()
        self.encoder = nn.LSTM(image_features, hidden, num_layers=1, bidirectional=True, batch_first=True)
        self.decoder = nn.LSTM(2 * hidden, hidden, num_layers=1, bidirectional=True, batch_first=True)
        self.linear_1 = nn.Linear(2 * hidden * window, 512)
        self.linear_2 = nn.Linear(512, 512)
        self.linear_3 = nn.Linear(512, 512)
        self.linear_4 = nn.Linear(512, 512)
        self.linear_5 = nn.Linear(512, 512)
        self.output_layer_type = nn.Linear(512, num_classes_type)
        self.activation = nn.SELU()

    def forward(self, x, return_hidden=False):
        x, _ = self.encoder(x)                    # zero initial state (simple_model.py:51)
        x, _ = self.decoder(x)                    # zero initial state again (:54)
        hid = x
        x = torch.flatten(x, start_dim=1, end_dim=2)
        for lin in (self.linear_1, self.linear_2, self.linear_3, self.linear_4, self.linear_5):
            x = self.activation(lin(x))           # dropout layers are identity in eval()
        p = torch.softmax(self.output_layer_type(x), dim=1)
        return (p, hid) if return_hidden else p


class PolishNet(nn.Module):
    def __init__(self, image_features=10, hidden=128, num_classes=5):
        super().__init__()
        self.gru_encoder = nn.GRU(image_features, hidden, num_layers=1, bidirectional=True, batch_first=True)
        self.gru_decoder = nn.GRU(2 * hidden, hidden, num_layers=1, bidirectional=True, batch_first=True)
        self.dense1 = nn.Linear(2 * hidden, num_classes)

    def forward(self, x, hidden):
        hidden = hidden.transpose(0, 1).contiguous()        # [B,2,H] -> [2,B,H]  (simple_model.py:28)
        x_out, hidden_out = self.gru_encoder(x, hidden)
        x_out, hidden_final = self.gru_decoder(x_out, hidden_out)
        return self.dense1(x_out), hidden_final.transpose(0, 1).contiguous()


def make_variant_weights(seed: int = 0, out_scale: float = 3.0) -> dict:
    """Seeded random weights (no trained checkpoint is available offline): default PyTorch init with
    the output layer scaled so that the softmax is not near-uniform."""
    torch.manual_seed(seed)
    m = VariantNet()
    with torch.no_grad():
        m.output_layer_type.weight.mul_(out_scale * 8)
        for lin in (m.linear_1, m.linear_2, m.linear_3, m.linear_4, m.linear_5):
            lin.weight.mul_(1.5)
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


def make_polish_weights(seed: int = 0, out_scale: float = 3.0) -> dict:
    torch.manual_seed(seed)
    m = PolishNet()
    with torch.no_grad():
        m.dense1.weight.mul_(out_scale * 4)
        for name, p in m.named_parameters():
            if "gru" in name and "weight" in name:
                p.mul_(2.0)
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


@torch.no_grad()
def variant_predict(state: dict, images_i8: np.ndarray, batch: int = 512, threads: int = 1, return_hidden: bool = False):
    """images int8 [N,33,26] -> float32 probs [N,3]  (my_collate casts to float, dataloader_predict.py:82-91)."""
    torch.set_num_threads(threads)
    m = VariantNet()
    m.load_state_dict(state)
    m.eval()
    outs, hids = [], []
    for b0 in range(0, images_i8.shape[0], batch):
        x = torch.from_numpy(images_i8[b0:b0 + batch].astype(np.float32))
        if return_hidden:
            p, h = m(x, return_hidden=True)
            hids.append(h.numpy())
        else:
            p = m(x)
        outs.append(p.numpy())
    probs = np.concatenate(outs) if outs else np.zeros((0, 3), np.float32)
    if return_hidden:
        return probs, (np.concatenate(hids) if hids else np.zeros((0, 33, 512), np.float32))
    return probs


@torch.no_grad()
def polish_predict(state: dict, images_u8: np.ndarray, batch: int = 128, threads: int = 1):
    """images uint8 [N,1000,10] -> bases uint8 [N,1000], phred uint8 [N,1000], hidden float32 [19,N,2,128]
    (the state returned by each window), acc float32 [N,1000,5]."""
    torch.set_num_threads(threads)
    m = PolishNet()
    m.load_state_dict(state)
    m.eval()
    N = images_u8.shape[0]
    bases = np.zeros((N, 1000), np.uint8)
    phred = np.zeros((N, 1000), np.uint8)
    hid_all = np.zeros((19, N, 2, 128), np.float32)
    acc_all = np.zeros((N, 1000, 5), np.float32)
    for b0 in range(0, N, batch):
        images = torch.from_numpy(images_u8[b0:b0 + batch]).type(torch.FloatTensor)       # cpu.py:51
        hidden = torch.zeros(images.size(0), 2, 128)                                       # :53
        acc = torch.zeros((images.size(0), images.size(1), 5))                             # :55
        w = 0
        for i in range(0, 1000, 50):                                                        # :57
            if i + 100 > 1000:
                break
            out, hidden = m(images[:, i:i + 100], hidden)                                   # :68
            hid_all[w, b0:b0 + images.size(0)] = hidden.numpy()
            w += 1
            pad = nn.ZeroPad2d((0, 0, i, 1000 - (i + 100)))
            acc = torch.add(acc, pad(torch.softmax(out, dim=2)))                            # :75-81
        values, labels = torch.max(acc, 2)                                                  # :83
        counts = torch.ones((values.size(0), values.size(1) - 100))
        counts = nn.ZeroPad2d((50, 50))(counts) + 1                                         # :86-88
        ph = -10 * torch.log10(1.0 - (values / counts))
        ph[ph == float("inf")] = 100
        bases[b0:b0 + images.size(0)] = labels.numpy().astype(np.uint8)
        phred[b0:b0 + images.size(0)] = ph.numpy().astype(np.uint8)                         # DataStorePredict.py:49
        acc_all[b0:b0 + images.size(0)] = acc.numpy()
    return bases, phred, hid_all, acc_all
