"""Recurrent roll-out that drives the cell: the reference's STC_Encoder (layers x T) followed by the
STC_Decoder horizon loop (/root/reference/framework/STC_GNN.py:97-123, 154-166, 194-204) with the
supports handed in -- i.e. STCGNN.forward minus MGP_Gen and out_proj.  This is the unit bench.py times
("one sample = one [T,N,C] window through all layers and timesteps") and the only way to drive
configurations whose N makes the reference's MixedFusion (N^2 x N^2 weights) unconstructible.
"""
from __future__ import annotations

import torch
from torch import nn

from .cell import STC_Cell


class RecurrentStack(nn.Module):
    def __init__(self, num_nodes: int, num_categories: int, Ks: int, Kc: int, input_dim: int, hidden_dim: int,
                 num_layers: int, out_horizon: int, use_bias=True, activation=None):
        super().__init__()
        self.num_layers, self.out_horizon, self.hidden_dim = num_layers, out_horizon, hidden_dim
        # construction order = the reference's (encoder cells, then decoder cells) for seeded-init parity
        self.encoder = nn.ModuleList(
            STC_Cell(num_nodes, num_categories, Ks, Kc, input_dim if i == 0 else hidden_dim, hidden_dim,
                     use_bias=use_bias, activation=activation) for i in range(num_layers))
        self.decoder = nn.ModuleList(
            STC_Cell(num_nodes, num_categories, Ks, Kc, hidden_dim, hidden_dim, use_bias=use_bias,
                     activation=activation) for _ in range(num_layers))

    def forward(self, Gs, Gc: torch.Tensor, X_seq: torch.Tensor) -> torch.Tensor:
        """X_seq [B,T,N,C,Din] -> decoder hidden states [B,horizon,N,C,h] (out_horizon = 0: the last encoder
        layer's hidden states [B,T,N,C,h])."""
        assert X_seq.dim() == 5, "X_seq must be [B,T,N,C,Din]"
        B, T = X_seq.shape[0], X_seq.shape[1]
        # Layer l > 0 consumes layer l-1's per-step outputs directly: the reference stacks them and slices the stack
        # again (STC_GNN.py:114,111), which in backward costs one zero-filled [B,T,N,C,h] tensor plus one full-size
        # add per timestep (select_backward) -- same values, none of that traffic.
        seq = [X_seq[:, t] for t in range(T)]
        last = []
        for cell in self.encoder:
            Ht = cell.init_hidden(B)
            outs = []
            for t in range(T):
                Ht = cell(Gs=Gs, Gc=Gc, Xt=seq[t], Ht_1=Ht)
                outs.append(Ht)
            seq = outs
            last.append(Ht)
        if self.out_horizon == 0:   # encoder only (the large synthetic graphs drive STC_Encoder directly)
            return torch.stack(seq, dim=1)
        states, x, outs = last, last[-1], []
        for _ in range(self.out_horizon):
            new_states, inp = [], x
            for l, cell in enumerate(self.decoder):
                inp = cell(Gs=Gs, Gc=Gc, Xt=inp, Ht_1=states[l])
                new_states.append(inp)
            states, x = new_states, new_states[-1]
            outs.append(x)
        return torch.stack(outs, dim=1)
