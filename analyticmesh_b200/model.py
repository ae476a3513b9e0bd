"""ReLU MLP container with additive skip connections.

API-compatible with the reference's ``MLP`` (reference backend/model.py:5-148): same constructor
arguments, same parameter names (``linears``, ``tms``), same ``forward(x, requires_outputs_list)``
contract and the same ``get_info()`` dictionary, so that user scripts and ONNX files written for
the reference keep working.  The code is written from the semantics, not from the source.

Architecture encoding (reference backend/model.py:20-36):
  nodes        [3, n_1, ..., n_D, 1]
  arc_table    one row per hidden layer h = 1..D; row h-1 = [k, src_0, tm_0, ..., src_{k-1}, tm_{k-1}]
               describes the k additive skips that enter the *pre-activation of fully-connected
               layer h* (the layer that consumes hidden layer h and produces hidden layer h+1, or
               the output when h = D).  src = 0 is the raw input, src = j >= 1 is the post-ReLU
               output of hidden layer j.
  arc_tm_shape shapes (out, in) of the skip transforms; [0, 0] means identity.
"""
import math

import torch


class MLP(torch.nn.Module):
    def __init__(self, nodes, arc_table, arc_tm_shape, initialization='geometric',
                 geometric_radius=1.0, enable_print=True):
        super().__init__()
        self.nodes = list(nodes)
        self.arc_table = [list(r) for r in arc_table]
        self.arc_tm_shape = [list(s) for s in arc_tm_shape]
        self.geometric_radius = geometric_radius
        assert len(self.arc_table) == len(self.nodes) - 2, "arc_table needs one row per hidden layer"
        for row in self.arc_table:
            assert len(row) == 1 + 2 * row[0], f"malformed arc_table row {row}"

        self.num_of_linears = len(self.nodes) - 1
        self.linears = torch.nn.ModuleList(
            torch.nn.Linear(self.nodes[i], self.nodes[i + 1], bias=True) for i in range(self.num_of_linears))
        if initialization == 'kaiming':
            for lin in self.linears:
                torch.nn.init.kaiming_normal_(lin.weight)
                torch.nn.init.zeros_(lin.bias)
        elif initialization == 'geometric':
            # SAL geometric initialisation (reference backend/model.py:63-70): the network starts as
            # an approximate signed distance to a sphere of radius `geometric_radius`.
            last = self.num_of_linears - 1
            for i, lin in enumerate(self.linears):
                if i != last:
                    torch.nn.init.normal_(lin.weight, 0.0, math.sqrt(2) / math.sqrt(self.nodes[i + 1]))
                    torch.nn.init.zeros_(lin.bias)
                else:
                    torch.nn.init.constant_(lin.weight, math.sqrt(math.pi) / math.sqrt(self.nodes[i]))
                    torch.nn.init.constant_(lin.bias, -geometric_radius)
        elif initialization is None:
            if enable_print:
                print("Warning: Initialization strategy for MLP is not specified.")
        else:
            raise Exception(f'Error: No such initialization: {initialization}')

        self.num_of_acts = self.num_of_linears
        self.acts = torch.nn.ModuleList(
            torch.nn.ReLU(inplace=True) if i != self.num_of_acts - 1 else torch.nn.Identity()
            for i in range(self.num_of_acts))

        self.num_of_tms = len(self.arc_tm_shape)
        self.tms = torch.nn.ModuleList(
            torch.nn.Identity() if (s[0] == 0 and s[1] == 0)
            else torch.nn.Linear(in_features=s[1], out_features=s[0], bias=False)
            for s in self.arc_tm_shape)

        self._skip_sources = sorted({row[1 + 2 * j] for row in self.arc_table for j in range(row[0])})
        self.outputs_list = []

    def forward(self, x, requires_outputs_list=False):
        if requires_outputs_list:
            self.outputs_list.clear()
        taps = {}
        last = self.num_of_linears - 1
        for i in range(self.num_of_linears):
            if i in self._skip_sources:
                taps[i] = x
            x = self.linears[i](x)
            if i >= 1:
                row = self.arc_table[i - 1]
                for j in range(row[0]):
                    x = x + self.tms[row[2 + 2 * j]](taps[row[1 + 2 * j]])
            x = self.acts[i](x)
            if i != last and requires_outputs_list:
                self.outputs_list.append(x)
        return x

    def get_info(self):
        """weights / biases / arc_tm (zeros([0,0]) for identity) / arc_table (int32, zero padded);
        same keys as reference backend/model.py:125-148."""
        weights = [lin.weight for lin in self.linears]
        biases = [lin.bias for lin in self.linears]
        arc_tm = [torch.zeros([0, 0]) if isinstance(tm, torch.nn.Identity) else tm.weight for tm in self.tms]
        width = max(len(r) for r in self.arc_table)
        table = torch.zeros([len(self.arc_table), width], dtype=torch.int32)
        for r, row in enumerate(self.arc_table):
            table[r, :len(row)] = torch.tensor(row, dtype=torch.int32)
        return {'weights': weights, 'biases': biases, 'arc_tm': arc_tm, 'arc_table': table}
