"""Measurement harness for the third metric (train steps/s): the CALLERS of the hot path, restated in
torch so that a step can be driven end to end without TensorFlow.  Not product code -- the dense
autoencoder stays in the host framework (cuBLAS through torch.nn.Linear); every hot op comes from
libemk through ``encodermap_b200``.

* ``EncoderMapStep``  -- what ``SequentialModel.train_step`` does (reference encodermap/models/models.py:3367-3401):
  encoder (sin/cos for periodic input) -> latent -> decoder (atan2) ; losses auto (mean |periodic distance|) +
  L2 regularisation + center + 500 x distance_loss (second encoder pass, loss_functions.py:277) ; Adam(1e-3, clipvalue=1).
* ``ADCStep``         -- what ``ADCFunctionalModel.train_step`` does (:2260-2521) with use_backbone_angles=True:
  PeriodicInput on angles+dihedrals -> encoder -> latent -> decoder -> angles / dihedrals ; BackMapLayer ->
  PairwiseDistances(C-alpha) ; losses dihedral + angle + cartesian (mean |pair distance difference|) +
  cartesian_distance_loss + center + regularisation.
"""
from __future__ import annotations

import math
import sys
from pathlib import Path

import torch
from torch import nn

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from encodermap_b200 import ADCParameters, Parameters, parallel  # noqa: E402
from encodermap_b200.graph import graphed_train_step  # noqa: E402
from encodermap_b200.loss_functions import cartesian_distance_loss, distance_loss  # noqa: E402
from encodermap_b200.loss_functions.loss_functions import cartesian_distance_loss_from_coordinates, fused_cartesian_loss  # noqa: E402
from encodermap_b200.misc.distances import periodic_distance  # noqa: E402
from encodermap_b200.models.layers import BackMapLayer, PairwiseDistances, PeriodicInput  # noqa: E402


def mlp(sizes, final_activation=False):
    layers = []
    for i, (a, b) in enumerate(zip(sizes[:-1], sizes[1:])):
        layers.append(nn.Linear(a, b))
        if i < len(sizes) - 2 or final_activation:
            layers.append(nn.Tanh())
    return nn.Sequential(*layers)


class EncoderMapStep(nn.Module):
    def __init__(self, input_dim: int, p: Parameters, dp_group=None):
        """dp_group: data-parallel training -- every rank holds its rows of the global batch, the distance cost is the
        cost of the GLOBAL batch (parallel.data_parallel_sigmoid_cost), all other terms are local means."""
        super().__init__()
        self.p = p
        self.periodic = p.periodicity < float("inf")
        d_in = input_dim * (2 if self.periodic else 1)
        self.periodic_input = PeriodicInput(p, "input")
        self.encoder_model = mlp([d_in, *p.n_neurons])
        self.decoder_model = mlp([p.n_neurons[-1], *p.n_neurons[-2::-1], d_in])
        self.dist_loss = distance_loss(self, p, process_group=dp_group, data_parallel=dp_group is not None)

    def encoder(self, x, training=False):
        if self.periodic:
            x = self.periodic_input(x)
        return self.encoder_model(x)

    def decoder(self, z):
        y = self.decoder_model(z)
        if self.periodic:
            s, c = torch.chunk(y, 2, dim=1)
            y = torch.atan2(s, c)
            if self.p.periodicity != 2 * math.pi:
                y = y / (2 * math.pi) * self.p.periodicity
        return y

    def loss(self, x):
        z = self.encoder(x)
        y = self.decoder(z)
        if self.periodic:
            auto = periodic_distance(x, y, self.p.periodicity).mean()
        else:
            auto = (x - y).abs().mean()
        reg = sum((m.weight ** 2).sum() for m in self.modules() if isinstance(m, nn.Linear)) * self.p.l2_reg_constant
        center = (z ** 2).mean() * self.p.center_cost_scale
        return self.p.auto_cost_scale * auto + reg + center + self.dist_loss(x)


class ADCStep(nn.Module):
    def __init__(self, n_atoms: int, p: ADCParameters, dp_group=None, fused_cartesian: bool = False):
        """fused_cartesian: PairwiseDistances("output") + cartesian_loss as one launch (loss_functions.fused_cartesian_loss)
        instead of the reference's two layers + elementwise loss."""
        super().__init__()
        self.fused_cartesian = fused_cartesian
        self.p = p
        self.n = n_atoms
        d_in = 2 * ((n_atoms - 2) + (n_atoms - 3))
        self.pi_angles = PeriodicInput(p, "angles")
        self.pi_dihedrals = PeriodicInput(p, "dihedrals")
        self.encoder_model = mlp([d_in, *p.n_neurons])
        self.decoder_model = mlp([p.n_neurons[-1], *p.n_neurons[-2::-1], d_in])
        self.backmap = BackMapLayer(n_atoms // 2 - 1, (n_atoms - 3) // 2)
        self.pairwise = PairwiseDistances(p, "pairwise")
        self.cart_dist_loss = cartesian_distance_loss(self, p, process_group=dp_group, data_parallel=dp_group is not None)
        self.fused_loss = fused_cartesian_loss(None, None, p)
        # fused mode also takes cartesian_distance_loss straight from the input coordinates: with both, the
        # step holds no (batch, n_pairs) tensor at all
        self.cart_dist_loss_xyz = cartesian_distance_loss_from_coordinates(self, p)
        self.dp_group = dp_group

    def encoder(self, inputs, training=False):
        angles, dihedrals = inputs[:2]
        return self.encoder_model(torch.cat([self.pi_angles(angles), self.pi_dihedrals(dihedrals)], dim=1))

    def loss(self, angles, dihedrals, cartesians, distances):
        z = self.encoder((angles, dihedrals))
        y = self.decoder_model(z)
        na, nd = self.n - 2, self.n - 3
        sa, sd, ca, cd = torch.split(y, [na, nd, na, nd], dim=1)
        out_angles, out_dihedrals = torch.atan2(sa, ca), torch.atan2(sd, cd)
        if self.dp_group is not None:
            # BackMapLayer uses the mean bond lengths of the BATCH (layers.py:970): the global batch in data-parallel training
            lengths = distances.mean(dim=0, keepdim=True)
            torch.distributed.all_reduce(lengths, group=self.dp_group)
            distances = (lengths / torch.distributed.get_world_size(self.dp_group)).expand_as(distances)
        back = self.backmap((distances, out_angles, out_dihedrals))
        fused_all = self.fused_cartesian and self.dp_group is None
        inp_pair = None if fused_all else self.pairwise(cartesians)
        dihedral_loss = periodic_distance(dihedrals, out_dihedrals, self.p.periodicity).mean()
        angle_loss = periodic_distance(angles, out_angles, self.p.periodicity).mean()
        if self.fused_cartesian:
            cartesian_loss = self.fused_loss(cartesians, back)
        else:
            out_pair = self.pairwise(back)
            cartesian_loss = (inp_pair - out_pair).abs().mean()
        reg = sum((m.weight ** 2).sum() for m in self.modules() if isinstance(m, nn.Linear)) * self.p.l2_reg_constant
        center = (z ** 2).mean() * self.p.center_cost_scale
        cart_dist = self.cart_dist_loss_xyz(cartesians, z) if fused_all else self.cart_dist_loss(inp_pair, z)
        return dihedral_loss + angle_loss + cartesian_loss + cart_dist + center + reg


class ADCSidechainStep(nn.Module):
    """The ADC model with reconstructed side chains (reference models/models.py:704-760, 935-944): encoder over the periodic
    embeddings of backbone angles, backbone dihedrals and side-chain dihedrals; decoder back to those plus the side-chain angles;
    BackMapLayerWithSidechains on (input backbone distances, decoded angles, decoded dihedrals, input side distances, decoded
    side angles, decoded side dihedrals); PairwiseDistances on the gathered atoms of input and output; the ADC losses."""

    def __init__(self, counts, p: ADCParameters):
        super().__init__()
        from encodermap_b200.models.layers import BackMapLayerWithSidechains

        self.p = p
        fd = {-1: {k + 1: int(c) for k, c in enumerate(counts)}}
        p.reconstruct_sidechains, p.sidechain_info = True, fd
        self.backmap = BackMapLayerWithSidechains(fd)
        n_res = len(counts)
        self.sizes = [3 * n_res - 2, 3 * n_res - 3, self.backmap.n_sidechains, sum(counts)]   # ca, cdih, sa, sdih
        d_in = 2 * (self.sizes[0] + self.sizes[1] + self.sizes[3])
        self.pi = PeriodicInput(p, "inputs")
        self.encoder_model = mlp([d_in, *p.n_neurons])
        self.decoder_model = mlp([p.n_neurons[-1], *p.n_neurons[-2::-1], 2 * sum(self.sizes)])
        self.pairwise = PairwiseDistances(p, "pairwise")
        self.cart_dist_loss = cartesian_distance_loss(self, p)

    def encoder(self, inputs, training=False):
        ca, cdih, sdih = inputs
        return self.encoder_model(torch.cat([self.pi(ca), self.pi(cdih), self.pi(sdih)], dim=1))

    def loss(self, ca, cdih, sa, sdih, cartesians, cd, sd):
        z = self.encoder((ca, cdih, sdih))
        y = self.decoder_model(z)
        sin, cos = torch.split(y, [sum(self.sizes)] * 2, dim=1)
        o_ca, o_cdih, o_sa, o_sdih = torch.split(torch.atan2(sin, cos), self.sizes, dim=1)
        back = self.backmap((cd, o_ca, o_cdih, sd, o_sa, o_sdih))
        inp_pair = self.pairwise(cartesians)
        out_pair = self.pairwise(back)
        per = self.p.periodicity
        dihedral_loss = periodic_distance(cdih, o_cdih, per).mean() + periodic_distance(sdih, o_sdih, per).mean()
        angle_loss = periodic_distance(ca, o_ca, per).mean() + periodic_distance(sa, o_sa, per).mean()
        cartesian_loss = (inp_pair - out_pair).abs().mean()
        reg = sum((m.weight ** 2).sum() for m in self.modules() if isinstance(m, nn.Linear)) * self.p.l2_reg_constant
        center = (z ** 2).mean() * self.p.center_cost_scale
        return dihedral_loss + angle_loss + cartesian_loss + self.cart_dist_loss(inp_pair, z) + center + reg


def time_steps(model, batch_fn, steps=20, warmup=5, graph=False, grad_sync=None):
    """steps/s of a full training step (forward, backward, [data-parallel gradient averaging], clip, Adam).  graph=True
    captures the whole step in one CUDA graph (encodermap_b200.graph.graphed_train_step) and replays it."""
    # fused multi-tensor Adam: one launch for all parameters (the default foreach path is ~14 launches per step, a quarter of
    # the kernel time of an ADC step next to the hot ops; the optimiser stays the framework's, as in the reference)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, capturable=graph, fused=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    params = [p for p in model.parameters() if p.requires_grad]

    def step(batch):
        opt.zero_grad(set_to_none=True)
        loss = model.loss(*batch)
        loss.backward()
        if grad_sync is not None:
            grad_sync(params)
        torch.nn.utils.clip_grad_value_(params, 1.0)
        opt.step()
        return loss.detach()

    if not graph:
        losses = []
        for it in range(warmup + steps):
            if it == warmup:
                torch.cuda.synchronize()
                e0.record()
            losses.append(step(batch_fn(it)))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        return 1e3 / ms, ms, float(losses[warmup]), float(losses[-1])

    gstep = graphed_train_step(model, model.loss, opt, batch_fn(0), clip_value=1.0, grad_sync=grad_sync)
    first = None
    for it in range(warmup + steps):
        if it == warmup:
            torch.cuda.synchronize()
            e0.record()
        out = gstep(*batch_fn(it))
        if it == warmup:
            first = out.clone()
    e1.record()
    torch.cuda.synchronize()
    last = gstep.outputs.clone()
    ms = e0.elapsed_time(e1) / steps
    return 1e3 / ms, ms, float(first), float(last)


def run_all(dev, steps=20):
    out = {}
    g = torch.Generator(device=dev).manual_seed(7)
    n_data = 16384
    centres = (torch.rand(16, 1024, device=dev, generator=g) * 2 - 1) * math.pi
    data = centres[torch.randint(0, 16, (n_data,), device=dev, generator=g)] + 0.3 * torch.randn(n_data, 1024, device=dev, generator=g)
    data = torch.remainder(data + math.pi, 2 * math.pi) - math.pi
    pts = torch.rand(6000, 3, device=dev, generator=g)
    n, b = 300, 1024
    dist = 0.13 + 0.02 * torch.rand(b, n - 1, device=dev, generator=g)
    ang = 1.9 + 0.3 * torch.rand(b, n - 2, device=dev, generator=g)
    dih = (torch.rand(b, n - 3, device=dev, generator=g) * 2 - 1) * math.pi
    with torch.no_grad():
        cart = BackMapLayer(n // 2 - 1, (n - 3) // 2)((dist, ang, dih))
    cases = {
        # configs[1]: periodic 1024-dim angular features, batch 4096
        "train_cfg1_encodermap_periodic_1024d_batch4096":
            (lambda: EncoderMapStep(1024, Parameters()), lambda it: (data[(it * 4096) % n_data:(it * 4096) % n_data + 4096],)),
        # configs[0] shape: 3-d points -> 2-d latent, batch 256, non-periodic, cube sigmoid parameters
        "train_cfg0_cube_batch256":
            (lambda: EncoderMapStep(3, Parameters(periodicity=float("inf"), dist_sig_parameters=(0.2, 3, 6, 1, 2, 6))),
             lambda it: (pts[(it * 256) % 5632:(it * 256) % 5632 + 256],)),
        # configs[2]: ADC, 100-residue chain, batch 1024
        "train_cfg2_adc_100res_batch1024":
            (lambda: ADCStep(n, ADCParameters(cartesian_pwd_start=1, cartesian_pwd_step=3, use_backbone_angles=True)),
             lambda it: (ang, dih, cart, dist)),
        # same with the Cartesian branch fused
        "train_cfg2_adc_100res_batch1024_fused_cartesian":
            (lambda: ADCStep(n, ADCParameters(cartesian_pwd_start=1, cartesian_pwd_step=3, use_backbone_angles=True), fused_cartesian=True),
             lambda it: (ang, dih, cart, dist)),
        # the reference's default atom selection (cartesian_pwd_* = None: all 300 backbone atoms, 44 850 pair dims), fused branch
        "train_cfg2_adc_100res_batch1024_all_atoms_fused_cartesian":
            (lambda: ADCStep(n, ADCParameters(use_backbone_angles=True), fused_cartesian=True), lambda it: (ang, dih, cart, dist)),
    }
    # the same model with reconstructed side chains on a ubiquitin-sized topology (76 residues, 448 atoms), at the
    # reference's default batch size and at 1024
    import numpy as np

    rs = np.random.default_rng(5)
    counts = [int(c) for c in rs.integers(0, 5, size=76)]
    counts[0], counts[-1] = 3, 0
    n_side, n_sdih = sum(c + 1 for c in counts if c > 0), sum(counts)
    for bs in (256, 1024):
        s_in = [1.9 + 0.3 * torch.rand(bs, 226, device=dev, generator=g), (torch.rand(bs, 225, device=dev, generator=g) * 2 - 1) * math.pi,
                1.85 + 0.3 * torch.rand(bs, n_side, device=dev, generator=g), (torch.rand(bs, n_sdih, device=dev, generator=g) * 2 - 1) * math.pi]
        s_cd = 0.13 + 0.02 * torch.rand(bs, 227, device=dev, generator=g)
        s_sd = 0.14 + 0.03 * torch.rand(bs, n_side, device=dev, generator=g)
        with torch.no_grad():
            from encodermap_b200.models.layers import BackMapLayerWithSidechains

            s_cart = BackMapLayerWithSidechains({-1: {k + 1: c for k, c in enumerate(counts)}})((s_cd, s_in[0], s_in[1], s_sd, s_in[2], s_in[3]))
        batch = (s_in[0], s_in[1], s_in[2], s_in[3], s_cart, s_cd, s_sd)
        cases[f"train_adc_sidechains_76res_batch{bs}"] = (
            lambda: ADCSidechainStep(counts, ADCParameters(cartesian_pwd_start=1, cartesian_pwd_step=3, use_backbone_angles=True, use_sidechains=True)),
            (lambda bt: (lambda it: bt))(batch))
    for name, (make, batch_fn) in cases.items():
        res = {}
        for mode in ("eager", "cuda_graph"):
            torch.manual_seed(0)
            m = make().to(dev)
            sps, ms, l0, l1 = time_steps(m, batch_fn, steps, graph=(mode == "cuda_graph"))
            res[mode] = {"steps_per_s": sps, "ms_per_step": ms, "loss_first": l0, "loss_last": l1}
        out[name] = res
    return out


if __name__ == "__main__":
    import json

    print(json.dumps(run_all(torch.device("cuda:0")), indent=1))


def run_dp(dev, group, steps=20, batch_global=4096, weak=False):
    """configs[1] as DATA-PARALLEL training over the ranks of `group`: every rank owns batch/G rows
    (strong scaling: global batch fixed) or `batch_global` rows (weak scaling: per-GPU batch fixed); the distance cost is the
    cost of the global batch (all-gather rows -> this rank's pair-tile slice -> all-reduce loss + reduce-scatter dL/dz), the
    dense-layer gradients are averaged with one flat all-reduce.  Returns steps/s (max over ranks), eager and graph-replayed."""
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    rows = batch_global if weak else batch_global // world
    g = torch.Generator(device=dev).manual_seed(7)   # same data set on every rank; each rank reads its own rows
    n_data = 4 * rows * world
    centres = (torch.rand(16, 1024, device=dev, generator=g) * 2 - 1) * math.pi
    data = centres[torch.randint(0, 16, (n_data,), device=dev, generator=g)] + 0.3 * torch.randn(n_data, 1024, device=dev, generator=g)
    data = torch.remainder(data + math.pi, 2 * math.pi) - math.pi

    def batch_fn(it):
        start = ((it % 4) * world + rank) * rows
        return (data[start:start + rows],)

    res = {"rows_per_gpu": rows, "global_batch": rows * world, "scaling": "weak" if weak else "strong"}
    for mode in ("eager", "cuda_graph"):
        torch.manual_seed(0)   # identical initial weights on every rank
        model = EncoderMapStep(1024, Parameters(), dp_group=group).to(dev)
        try:
            sps, ms, l0, l1 = time_steps(model, batch_fn, steps, graph=(mode == "cuda_graph"),
                                         grad_sync=lambda params: parallel.average_gradients(params, group))
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
            res[mode] = {"steps_per_s": 1e3 / t.item(), "ms_per_step": t.item(), "loss_first": l0, "loss_last": l1}
        except Exception as e:  # noqa: BLE001 - reported, the other mode still runs
            res[mode] = {"error": repr(e)[:300]}
            torch.cuda.synchronize()
    return res
