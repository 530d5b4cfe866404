"""Batch-sharded graph evaluation (north star item 5).

The reference evaluates one graph in one thread on one host and has no
communication layer.  Graphs whose leaves carry a batch
dimension and whose root is a sum over it shard naturally: in ade order the
numpy leading (batch) dimension is the slowest one (llo/python/llo.cpp:18-22),
so a shard is a contiguous slice of every batched leaf.  One process drives one
GPU; every rank

  1. builds the same graph over its slice of the batched leaves (shapes are
     static in ade::Functor::get, so the graph is rebuilt with batch / G rows),
  2. derives the gradients once (llo::derive, llo/src/zprune.cpp:109-116) and
     evaluates all of them in ONE plan, sharing the forward pass,
  3. packs them into one flat device buffer and folds it over the ranks with a
     single NCCL allreduce over NVLink 5 / NVSwitch (llo_cuda_allreduce).

Scaling the loss by the GLOBAL batch makes the sum of the shard gradients the
full-batch gradient (up to fp reassociation across the G partial sums).

The host-side logic (slicing, packing, ordering) is independent of the device:
`ShardPlan` and `GradientPack` are plain Python/numpy and are what the CPU
(gloo, world_size 2) tests drive through a host backend.  The product backend
is `DeviceBackend`; there is no CPU fallback in this module.
"""
import os

import numpy as np


class ShardPlan:
    """Contiguous split of `batch` rows over `world` ranks; the remainder goes
    to the first ranks one row each."""

    def __init__(self, batch, world):
        if world < 1 or batch < world:
            raise ValueError("cannot shard %d rows over %d ranks" % (batch, world))
        self.batch = int(batch)
        self.world = int(world)

    def bounds(self, rank):
        if not 0 <= rank < self.world:
            raise ValueError("rank %d outside [0, %d)" % (rank, self.world))
        base, extra = divmod(self.batch, self.world)
        lo = rank * base + min(rank, extra)
        return lo, lo + base + (1 if rank < extra else 0)

    def rows(self, rank):
        lo, hi = self.bounds(rank)
        return hi - lo

    def slice(self, array, rank):
        """rows of a batched (leading-dimension) array owned by `rank`"""
        if array.shape[0] != self.batch:
            raise ValueError("array has %d rows, the plan shards %d" % (array.shape[0], self.batch))
        lo, hi = self.bounds(rank)
        return array[lo:hi]


class GradientPack:
    """Layout of several tensors in one flat buffer: offsets in elements, each
    rounded up to 16 bytes so every tensor stays vector-aligned."""

    def __init__(self, shapes, dtype):
        self.dtype = np.dtype(dtype)
        self.shapes = [tuple(int(d) for d in s) for s in shapes]
        align = max(1, 16 // self.dtype.itemsize)
        self.offsets = []
        total = 0
        for shape in self.shapes:
            self.offsets.append(total)
            count = int(np.prod(shape, dtype=np.int64)) if len(shape) else 1
            total += -(-count // align) * align
        self.count = total

    @property
    def nbytes(self):
        return self.count * self.dtype.itemsize

    def sizes(self):
        return [int(np.prod(s, dtype=np.int64)) if len(s) else 1 for s in self.shapes]

    def pack(self, arrays):
        flat = np.zeros(self.count, dtype=self.dtype)
        for off, size, a in zip(self.offsets, self.sizes(), arrays):
            flat[off:off + size] = np.asarray(a, dtype=self.dtype).reshape(-1)
        return flat

    def unpack(self, flat):
        return [flat[off:off + size].reshape(shape)
                for off, size, shape in zip(self.offsets, self.sizes(), self.shapes)]


class DeviceBackend:
    """Product backend: a binding of llo::ShardedEvaluator (csrc/host/llo/sharded.hpp).
    Deriving, the optimize passes, the packed layout, writing every gradient straight
    into its slot and the bucketed, overlapped NCCL fold all happen in C++; this class
    only keeps one evaluator per (loss, params) and hands numpy views back."""

    native = True

    def __init__(self, ordered=False, overlap=True, solo_rank=False):
        from . import llo
        if not llo.available():
            raise RuntimeError("batch-sharded evaluation needs a B200; there is no CPU fallback")
        self.llo = llo
        self.ordered = bool(ordered)
        # LLO_SHARD_OVERLAP=0: one fold after the last gradient (A/B runs)
        self.overlap = bool(overlap) and os.environ.get("LLO_SHARD_OVERLAP", "1") != "0"
        # LLO_SHARD_GRAPH=0: plan and launch every step instead of replaying a CUDA graph
        self.graph = os.environ.get("LLO_SHARD_GRAPH", "1") != "0"
        # LLO_SHARD_ORDER=backward: last layer's gradients (and their folds) first; the default
        # evaluates the first layer's gradients first, see llo/sharded.hpp deep_first_
        self.deep_first = os.environ.get("LLO_SHARD_ORDER", "deep") != "backward"
        self.solo_rank = bool(solo_rank)
        self._evaluators = {}

    def init_comm(self, rank, world, exchange):
        if self.solo_rank:
            if world != 1:
                raise ValueError("the solo backend evaluates on one rank")
            return
        if self.llo.comm_size() == world and self.llo.comm_rank() == rank and world > 1:
            return
        if world == 1:
            self.llo.comm_init(0, 1)
            return
        ident = self.llo.comm_unique_id() if rank == 0 else None
        ident = exchange(ident)
        self.llo.comm_init(rank, world, ident)

    def evaluator(self, loss, params, passes):
        key = (id(loss), tuple(id(p) for p in params), passes)
        if key not in self._evaluators:
            ev = self.llo.ShardedEvaluator(loss, list(params), passes, self.ordered, self.overlap,
                                           self.solo_rank, self.graph, self.deep_first)
            self._evaluators[key] = (loss, list(params), ev)     # keeps the ids alive
        return self._evaluators[key][2]

    def solo(self):
        """this rank alone on its own GPU, leaving the ranks' communicator untouched: the
        single-GPU evaluation a multi-rank run checks its agreed gradient against"""
        return DeviceBackend(ordered=False, overlap=False, solo_rank=True)

    def to_host(self, buf, pack):
        return pack.unpack(buf.numpy(pack.dtype, 0, pack.count))

    def scalar(self, buf, pack, index):
        return buf.numpy(pack.dtype, pack.offsets[index] * pack.dtype.itemsize, 1)[0]


def _default_exchange(ident):
    """broadcast rank 0's communicator id through torch.distributed"""
    import torch.distributed as dist
    if not dist.is_initialized():
        raise RuntimeError("no id exchange given and torch.distributed is not initialised")
    box = [ident]
    dist.broadcast_object_list(box, src=0)
    return box[0]


class BatchShardedExecutor:
    """Gradient evaluation of one graph sharded by batch over `world` ranks.

        ex = BatchShardedExecutor(batch=65536)              # rank / world from the env
        x = llo.variable(ex.shard(X), "X") ...              # this rank's rows
        loss = ... / ex.batch                                # scale by the GLOBAL batch
        grads = ex.grad_eval(loss, params, np.float32)       # list of numpy arrays, summed over ranks
    """

    def __init__(self, batch, rank=None, world=None, backend=None, exchange=None, ordered=False,
                 passes=None):
        self.rank = int(os.environ.get("RANK", "0")) if rank is None else int(rank)
        self.world = int(os.environ.get("WORLD_SIZE", "1")) if world is None else int(world)
        self.plan = ShardPlan(batch, self.world)
        self.batch = int(batch)
        self.backend = backend if backend is not None else DeviceBackend(ordered=ordered)
        self.backend.init_comm(self.rank, self.world, exchange or _default_exchange)
        self._derived = {}
        # graph -> graph passes applied to the derived gradients before they are
        # evaluated (llo.optimize; None = llo.OPT_ALL, 0 = evaluate as derived).
        # OPT_DIV_CANCEL is what turns the product rule's (a*b)/a into b, i.e. what
        # lets the matmul gradients lower to GEMMs at all; it is inexact by <= 1 ulp
        # per term and is named here, not hidden in the evaluator.
        self.passes = passes
        self.opt_report = None

    def shard(self, array):
        return np.ascontiguousarray(self.plan.slice(np.asarray(array), self.rank))

    def local_rows(self):
        return self.plan.rows(self.rank)

    def _native(self):
        return getattr(self.backend, "native", False)

    def _passes(self):
        from . import llo
        return llo.OPT_ALL if self.passes is None else int(self.passes)

    def derive(self, loss, params):
        """gradient graphs of `loss` w.r.t. every parameter (cached per loss)"""
        if self._native():
            ev = self.backend.evaluator(loss, params, self._passes())
            self.opt_report = ev.opt_report()
            return ev.roots()[1:]
        from . import llo
        key = (id(loss), tuple(id(p) for p in params))
        if key not in self._derived:
            grads = [llo.derive(loss, p) for p in params]
            passes = llo.OPT_ALL if self.passes is None else int(self.passes)
            if passes:
                # the loss is rewritten with the gradients so they keep sharing the forward pass
                out, self.opt_report = llo.optimize([loss] + grads, passes)
                opt_loss, grads = out[0], list(out[1:])
            else:
                opt_loss = loss
            self._derived[key] = (loss, list(params), grads, opt_loss)
        return self._derived[key][2]

    def grad_eval_device(self, loss, params, dtype, with_loss=False):
        """one gradient evaluation; returns (buffer, pack) with the summed
        gradients (and the summed loss first when with_loss) in HBM"""
        if self._native():
            ev = self.backend.evaluator(loss, params, self._passes())
            if self.opt_report is None:
                self.opt_report = ev.opt_report()
            shapes = [self._shape(p) for p in params]
            if with_loss:
                shapes = [()] + shapes
            return ev.grad_eval_device(np.dtype(dtype), with_loss), GradientPack(shapes, dtype)
        roots = list(self.derive(loss, params))
        shapes = [self._shape(p) for p in params]
        if with_loss:
            key = (id(loss), tuple(id(p) for p in params))
            roots = [self._derived[key][3]] + roots
            shapes = [()] + shapes
        pack = GradientPack(shapes, dtype)
        results = self.backend.evaluate_many(roots, dtype)
        return self.backend.reduce_packed(results, pack), pack

    def grad_eval(self, loss, params, dtype, with_loss=False):
        buf, pack = self.grad_eval_device(loss, params, dtype, with_loss)
        return self.backend.to_host(buf, pack)

    @staticmethod
    def _shape(tensor):
        return tuple(int(d) for d in tensor.shape())
