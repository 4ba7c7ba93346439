"""The optimiser loop (``ctdr/optimize.py:95-257``): Adam over ``displace`` and ``mus`` against the
masked-L2 data term plus edge / Laplacian / flatten regularisers, same ``log.txt`` line format.

``run(model, ds, niter, args, epoch_start=0, use_adam=True)`` keeps the reference signature.  Two
data paths: the autograd path (``model(idx_angles, wedge)`` like the reference) and, with
``fused=True`` (default when the model uses the fused projector), ``projection_step`` — one
device sweep that forms the residual, the loss and the gradients without the reference's three
host syncs per iteration (``rasterizer.py:64``, ``fp_vecs.py:18``, the boolean-mask index at
``optimize.py:168``) and all-reduces them when angles are sharded over several GPUs.
"""
from __future__ import annotations

import time

import numpy as np
import torch

from . import distributed
from .function import rasterizer as R
from .utils import util_mesh, util_vis


def get_params(model, exclude_mus=False):
    return model.parameters()


def projection_step(model, idx_angles, p_batch, exchange=None):
    """Data term of one iteration through the fused device path.

    Adds the gradient of ``mean_valid (p - phat)^2`` to ``.grad`` of every trainable parameter the displaced
    vertices depend on — ``displace``, and ``center`` / ``scale`` when the model registers them as parameters
    (``ctdr/model/vanilla.py:80-83,123-139``), exactly the leaves ``loss.backward()`` reaches in the reference
    (``ctdr/optimize.py:190``) — and to ``mus.grad``; returns ``(data_loss, phat, n_valid)`` as device tensors.
    The vertex cotangent comes out of the fused kernels; ``torch.autograd.backward`` pushes it through the few
    torch ops of ``get_displaced_vertices``.  With several ranks, each passes its own angle shard and the sums
    are all-reduced.
    """
    fp = model.fp
    with torch.enable_grad():
        v = model.get_displaced_vertices()
    verts = v.detach().contiguous()
    if model.mus_fixed_no == 1:
        model.mus.data[0] = model.p_min
    mus = model.mus.detach().contiguous()
    dev = verts.device
    if exchange is None:
        # ad-hoc buffers: no peer blocks (setting those up is a collective and worth it only for a persistent exchange)
        exchange = distributed.GradientExchange(verts.shape[0], mus.numel(), dev, peer=False)
    exchange.zero_()
    idx = R._device_idx(idx_angles, dev, fp.geom.nangles)
    phat, _, _, _ = R.project_residual_step(fp.geom, verts, fp._faces32(model.faces), fp.labels_fx2, mus, idx,
                                            p_batch, grad_verts=exchange.grad_verts,
                                            grad_mus=exchange.grad_mus, out2=exchange.scalars)
    exchange.all_reduce()
    gverts, gmus, loss = exchange.mean_gradients()
    if v.requires_grad:
        torch.autograd.backward(v, gverts.to(v.dtype))        # -> displace.grad (+ scale.grad, center.grad)
    if isinstance(model.mus, torch.nn.Parameter) and model.mus.requires_grad:
        model.mus.grad = gmus.clone() if model.mus.grad is None else model.mus.grad + gmus
    return loss.to(torch.float32), phat, exchange.scalars[1]


class GraphedIteration:
    """One full-batch optimiser iteration captured in a CUDA graph.

    The iteration of ``run`` — fused data term, sparse regularisers and their autograd, Adam step,
    ``mus.clamp_(min=0)`` — is launch-bound at the reference's own problem sizes (config 1/2: tens of
    microseconds of kernels behind ~100 launches).  Capturing it once and replaying removes the
    Python/launch overhead; nothing in the iteration synchronises with the host.  ``step()`` replays
    and returns the device tensor ``[data_loss, edge, lap, flat]`` of the iteration.
    """

    def __init__(self, model, args, idx_angles, p_full, lr=None):
        self.model, self.args = model, args
        dev = model.vertices.device
        self.idx = torch.as_tensor(idx_angles).to(device=dev, dtype=torch.int64)
        self.p = p_full.to(dev).contiguous()
        self.exchange = distributed.GradientExchange(model.vertices.shape[0], model.mus.numel(), dev)
        params = [p for p in model.parameters() if p.requires_grad]
        self.lr = torch.tensor(float(args.lr if lr is None else lr), device=dev)
        self.opt = torch.optim.Adam(params, lr=self.lr, betas=(0.9, 0.99), capturable=True)
        self.stats = torch.zeros(4, device=dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(3):                       # warm-up on a side stream (allocator, Adam state)
                self._iteration()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        self.opt.zero_grad(set_to_none=True)
        from . import _lib
        launches0 = _lib.lib().ctdr_launch_count()
        with torch.cuda.graph(self.graph):
            self._iteration()
        self.ctdr_launches_per_iteration = int(_lib.lib().ctdr_launch_count() - launches0)   # library kernels in the graph
        self.phat = self._phat

    def _iteration(self):
        model, args = self.model, self.args
        self.opt.zero_grad(set_to_none=True)
        data_loss, self._phat, _ = projection_step(model, self.idx, self.p, self.exchange)
        vertices = model.get_displaced_vertices()
        zero = torch.zeros((), device=vertices.device)
        edge = R_edge(model, vertices) if args.wedge > 0. else zero
        lap = model.laplacian_loss(vertices).mean() if model.use_lap_loss else zero
        flat = model.flat_loss(vertices) if model.use_flat_loss else zero
        reg = args.wedge * edge + args.wlap * lap + args.wflat * flat
        if reg.requires_grad:
            reg.backward()
        self.opt.step()
        model.mus.data.clamp_(min=0.0)
        self.stats.copy_(torch.stack([data_loss.detach(), edge.detach(), lap.detach(), flat.detach()]))

    def set_lr(self, lr):
        self.lr.fill_(float(lr))

    def step(self):
        self.graph.replay()
        return self.stats


def run(model, ds, niter, args, epoch_start=0, use_adam=True, fused=None, log_every=60, graphed=False,
        decay_lr=False):
    """``decay_lr``: the reference halves ``args.lr`` 100 iterations before the end of runs longer than 400
    (``ctdr/optimize.py:148-152``) but never hands the new value to the optimiser, so its Adam step size stays
    constant; the default reproduces exactly that (``args.lr`` changes, the optimiser does not).  ``decay_lr=True``
    applies the halving to the optimiser as the reference apparently intended."""
    if fused is None:
        fused = getattr(model.fp, "fused", False)
    print("@ model.mus", model.mus)
    params = list(get_params(model))
    opt = (torch.optim.Adam(params, args.lr, betas=(0.9, 0.99)) if use_adam
           else torch.optim.SGD(params, lr=args.lr, momentum=0.9, nesterov=True))
    f = open(args.dresult + "log.txt", "w" if epoch_start == 0 else "a")
    log = f"@ statistics of mesh: {model.vertices.shape[0]}, {model.faces.shape[0]}\n"
    dev = model.vertices.device
    if args.b == 0:
        ds_loader = [[torch.arange(ds.nangles), ds.p.to(dev)]]
    else:
        ds_loader = torch.utils.data.DataLoader(ds, batch_size=args.b, shuffle=True, num_workers=0,
                                                drop_last=True, pin_memory=True)
    exchange = distributed.GradientExchange(model.vertices.shape[0], model.mus.numel(), dev) if fused else None
    graph_it = None
    if graphed and fused and use_adam and args.b == 0:
        graph_it = GraphedIteration(model, args, ds_loader[0][0], ds_loader[0][1])
        opt = graph_it.opt
    ledge = llap = lflat = 0.
    loss_prev = None
    phat = mask_valid = None
    for epoch in range(epoch_start, niter):
        if epoch + 100 == niter and niter > 400:
            args.lr *= 0.5                                  # optimize.py:150-152
            print("@ args.lr", args.lr)
            if decay_lr:
                for group in opt.param_groups:
                    group["lr"] *= 0.5
                if graph_it is not None:
                    graph_it.set_lr(graph_it.lr.item() * 0.5)
        torch.cuda.synchronize(dev)
        start = time.time()
        if graph_it is not None:
            vals = graph_it.step().tolist()              # the one host sync of the iteration
            data_loss, edge_loss, lap_loss, flat_loss = (torch.tensor(v) for v in vals)
            loss = data_loss + args.wedge * edge_loss + args.wlap * lap_loss + args.wflat * flat_loss
            phat, mask_valid = graph_it.phat, None
        for idx_angles, p_batch in (ds_loader if graph_it is None else []):
            if args.b > 0:
                p_batch = p_batch.to(dev, non_blocking=True)
            opt.zero_grad()
            if fused:
                data_loss, phat, _ = projection_step(model, idx_angles, p_batch, exchange)
                vertices = model.get_displaced_vertices()
                edge_loss = R_edge(model, vertices) if args.wedge > 0. else 0.
                lap_loss = model.laplacian_loss(vertices).mean() if model.use_lap_loss else 0.
                flat_loss = model.flat_loss(vertices) if model.use_flat_loss else 0.
                reg = args.wedge * edge_loss + args.wlap * lap_loss + args.wflat * flat_loss
                if torch.is_tensor(reg):
                    reg.backward()                     # adds to the gradients set by projection_step
                loss = data_loss + (reg.detach() if torch.is_tensor(reg) else reg)
                mask_valid = None
            else:
                phat, mask_valid, edge_loss, lap_loss, flat_loss = model(idx_angles, args.wedge)
                data_loss = (p_batch - phat)[mask_valid].pow(2).mean()
                loss = data_loss + args.wedge * edge_loss + args.wlap * lap_loss + args.wflat * flat_loss
                loss.backward()
            opt.step()
            model.mus.data.clamp_(min=0.0)
        loss_now = float(loss.detach())               # the one host sync of the iteration
        elapsed = time.time() - start
        if loss_prev is not None and args.b == 0 and 0 < loss_prev - loss_now < 1e-11:
            print("! converged")
            break
        loss_prev = loss_now
        scalar = lambda v: float(v.detach()) if torch.is_tensor(v) else float(v)
        ledge = scalar(edge_loss) if args.wedge > 0. else ledge
        llap = scalar(lap_loss) if args.wlap > 0. else llap
        lflat = scalar(flat_loss) if args.wflat > 0. else lflat
        log += (f"~ {epoch:03d} l2_loss: {float(data_loss.detach()):.8f} edge: {ledge:.6f} lap: {llap:.6f} flat: {lflat:.6f} "
                f"mus: {str(model.mus.cpu().detach().numpy())} time: {elapsed:.4f}\n")
        if epoch % log_every == 0 or epoch == niter - 1:
            if mask_valid is None:
                from .nn.fp_vecs import get_valid_mask
                mask_valid = get_valid_mask(phat, model.fp.max_sino_val)
            if torch.sum(~mask_valid) > 15000 and epoch > 100:
                raise AssertionError("consider increasing regularization")
            print(log)
            if args.b == 0 and ds.p_original is not None:
                res_np = ds.p_original - phat.detach().cpu().numpy()
                f.write(f"~ res_np: {np.mean(res_np ** 2)}\n")
                util_vis.save_sino_as_img(args.dresult + f"{epoch:04d}_sino_res.png", res_np, cmap="coolwarm")
            if getattr(args, "verbose", 0) == 0:
                continue
            vv = (model.vertices + model.displace.detach()).cpu()
            util_vis.save_sino_as_img(args.dresult + f"{epoch:04d}_sino.png", phat.detach().cpu().numpy())
            util_mesh.save_mesh(args.dresult + f"{epoch:04d}.obj", vv.numpy(), model.faces.cpu().numpy(),
                                model.labels_v_np, model.labels.cpu().numpy())
            if epoch == niter - 1:
                util_mesh.save_mesh(args.dresult + "mesh.obj", vv.numpy(), model.faces.cpu().numpy(),
                                    model.labels_v_np, model.labels.cpu().numpy())
    f.write(log + "\n")
    f.close()
    return phat


def R_edge(model, vertices):
    from .utils import util_fun
    return util_fun.compute_edge_length(vertices, model.faces)
