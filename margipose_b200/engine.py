"""Static execution engine for the MargiPose network body on one B200.

Walks the module tree of `margipose_b200.models.margipose_model.MargiPoseModel` (which mirrors
/root/reference/src/margipose/models/margipose_model.py:153-200 and the truncated torchvision
ResNet of :119-138) once per (batch, resolution), allocates every activation / gradient buffer
up front (bf16 NHWC, channels padded to 64), and records the forward and backward passes as flat
lists of C-ABI launches with pre-built argument structs.  Running a pass is then a loop of ctypes
calls -- no allocation, no autograd graph, no host synchronisation -- which is also what makes
the whole training step capturable in one CUDA graph.  The three HeatmapColumns of a stage are
independent (margipose_model.py:196-198) and structurally identical: their programs are zipped into
ONE program of grouped launches (three problems per kernel launch); MARGIPOSE_B200_GROUP=0 runs
them on three streams instead.

Parameters live in ONE flat fp32 buffer (conv weights in channels-last memory = the GEMM's
[rows][taps][cols] order), gradients in a second flat buffer with the same layout, and the bf16
GEMM operands in a third that one mp_pack_weights launch refreshes.  The nn.Parameters of the
module tree are views into the flat buffers, so state_dict()/load_state_dict()/optimizers see
ordinary tensors with the reference's key names and shapes.
"""
import ctypes
import os

import torch
from torch import nn

from . import convops as C
from . import dsntnn as K
from ._lib import (BnArgs, BnBranch, BnFoldEntry, PackEntry, lib, check, stream_ptr, planes, MargiposeB200Error)


def _ptr(t):
    return t.data_ptr() if t is not None else None


# ------------------------------------------------------------------------------------ parameters
class _Slot:
    """A parameter (or buffer) of the module tree and its place in a flat buffer."""

    def __init__(self, mod, name, off, numel, master_shape=None, conv_k=None):
        self.mod, self.name, self.off, self.numel = mod, name, off, numel
        self.master_shape = master_shape     # conv weights: (rows, taps, cols)
        self.conv_k = conv_k
        self.data = None                     # flat view, master layout
        self.grad = None


class ParamBank:
    """Flat storage for parameters, gradients, BatchNorm buffers and bf16 weight packs."""

    def __init__(self, split=False):
        self.split = bool(split)       # bf16x3 precision mode: every weight pack has a low half (W_lo)
        self.pairs = C.SplitPairs() if split else None
        self.params, self.buffers, self.counters = [], [], []
        self.n_param = self.n_buffer = 0
        self.pack_entries = []          # dicts; turned into a device table by finalize()
        self.n_pack = self.n_work = 0
        self.pack_head = None           # (entries, work) of the ResNet stem's packs: pack(part=...) can refresh them apart
        self._seen = {}
        self.flat = self.flat_grad = self.flat_buf = self.flat_cnt = self.packs = None
        self._mats = []

    def _align(self, n, a=8):
        return (n + a - 1) // a * a

    def add_param(self, mod, name, master_shape=None, conv_k=None):
        key = (id(mod), name)
        if key in self._seen:
            return self._seen[key]
        p = getattr(mod, name)
        s = _Slot(mod, name, self.n_param, p.numel(), master_shape, conv_k)
        self.n_param += self._align(p.numel())
        self.params.append(s)
        self._seen[key] = s
        return s

    def add_buffer(self, mod, name):
        key = (id(mod), name)
        if key in self._seen:
            return self._seen[key]
        b = getattr(mod, name)
        if b.dtype == torch.int64:
            s = _Slot(mod, name, len(self.counters), 1)
            self.counters.append(s)
        else:
            s = _Slot(mod, name, self.n_buffer, b.numel())
            self.n_buffer += self._align(b.numel())
            self.buffers.append(s)
        self._seen[key] = s
        return s

    def add_matrix(self, rows_p, parts):
        """A bf16 GEMM operand [rows_p][sum of taps*cols_p]; parts = [(slot, transpose, A, B, taps,
        cols_p)] laid side by side along K.  Returns a handle with .t (tensor) after finalize()."""
        k_total = sum(taps * cols_p for (_s, _tr, _a, _b, taps, cols_p) in parts)
        # split (bf16x3) mode: every row holds the hi halves of all K columns, then their lo halves -- ONE matrix, so a
        # convolution's hi*hi, lo*hi and hi*lo products can share a launch (the W_lo taps read at K offset + k_total)
        row = k_total * (2 if self.split else 1)
        mat = type('Mat', (), {})()
        mat.rows, mat.k, mat.off, mat.koffs, mat.t, mat.row = rows_p, k_total, self.n_pack, [], None, row
        koff = 0
        for slot, transpose, a, b, taps, cols_p in parts:
            work = rows_p * taps * cols_p
            self.pack_entries.append(dict(slot=slot, dst_off=self.n_pack + koff, row_stride=row,
                                          work_off=self.n_work, work_end=self.n_work + work, A=a, B=b,
                                          taps=taps, transpose=int(transpose), rows_p=rows_p,
                                          cols_p=cols_p, lo_off=k_total if self.split else 0))
            self.n_work += work
            mat.koffs.append(koff)
            koff += taps * cols_p
        self.n_pack += rows_p * row
        self._mats.append(mat)
        return mat

    def mark_pack_head(self):
        """Everything packed so far belongs to the feature extractor (needed first in a forward pass)."""
        self.pack_head = (len(self.pack_entries), self.n_work)

    def finalize(self, device):
        self.flat = torch.zeros(max(self.n_param, 8), device=device)
        self.flat_grad = torch.zeros_like(self.flat)
        self.flat_buf = torch.zeros(max(self.n_buffer, 8), device=device)
        self.flat_cnt = torch.zeros(max(len(self.counters), 1), dtype=torch.int64, device=device)
        self.packs = torch.zeros(max(self.n_pack, 8), dtype=torch.bfloat16, device=device)
        for s in self.params:
            p = getattr(s.mod, s.name)
            s.data = self.flat[s.off:s.off + s.numel]
            s.grad = self.flat_grad[s.off:s.off + s.numel]
            if s.master_shape is not None:      # conv weight: channels-last memory, torch-shaped view
                rows, taps, cols = s.master_shape
                k = s.conv_k
                view = s.data.view(rows, k, k, -1).permute(0, 3, 1, 2)
                gview = s.grad.view(rows, k, k, -1).permute(0, 3, 1, 2)
                s.data = s.data.view(rows, taps, cols)
                s.grad = s.grad.view(rows, taps, cols)
            else:
                view, gview = s.data.view(p.shape), s.grad.view(p.shape)
            with torch.no_grad():
                view.copy_(p.detach().to(device=device, dtype=torch.float32))
            p.data = view
            p.grad = None
            s.view, s.gview, s.ptr = view, gview, view.data_ptr()
        for s in self.buffers:
            b = getattr(s.mod, s.name)
            s.data = self.flat_buf[s.off:s.off + s.numel]
            s.data.copy_(b.detach().to(device=device, dtype=torch.float32))
            s.mod._buffers[s.name] = s.data.view(b.shape)
        for s in self.counters:
            b = getattr(s.mod, s.name)
            self.flat_cnt[s.off] = int(b)
            s.mod._buffers[s.name] = self.flat_cnt[s.off]
        table = (PackEntry * max(len(self.pack_entries), 1))()
        head_n, head_work = self.pack_head if self.pack_head else (0, 0)
        for i, e in enumerate(self.pack_entries):
            t = table[i]
            t.src_off, t.dst_off, t.dst_row_stride = e['slot'].off, e['dst_off'], e['row_stride']
            # two tables in one array: [0, head_n) counts its work from 0, and so does [head_n, n)
            base = head_work if i >= head_n else 0
            t.work_off, t.work_end = e['work_off'] - base, e['work_end'] - base
            t.A, t.B, t.taps, t.transpose = e['A'], e['B'], e['taps'], e['transpose']
            t.rows_p, t.cols_p = e['rows_p'], e['cols_p']
            t.lo_off = e['lo_off']
        raw = torch.frombuffer(bytearray(bytes(table)), dtype=torch.uint8)
        self.pack_table = raw.to(device)
        for m in self._mats:
            m.t = self.packs[m.off:m.off + m.rows * m.row].view(m.rows, m.row)
        self.device = device

    def linked(self):
        """True while every nn.Parameter still aliases its slice of the flat buffer (a rebound parameter,
        e.g. `m.weight = nn.Parameter(...)` or `m.weight.data = pretrained`, makes the model re-materialise)."""
        for s in self.params:
            p = s.mod._parameters.get(s.name)
            if p is None or p.data_ptr() != s.ptr:
                return False
        return True

    def param_version(self):
        """Changes whenever any parameter is written in place through its nn.Parameter (load_state_dict,
        torch.optim steps, p.add_): the Parameters carry their own version counters -- `p.data = view` in
        finalize() detaches them from the flat buffer's -- so the bf16 weight-pack cache keys on their sum."""
        return self.flat._version + sum(s.mod._parameters[s.name]._version for s in self.params)

    def pack(self, part=None):
        """fp32 master weights -> bf16 GEMM operands.  part=None: everything; 0: the feature extractor's packs only;
        1: the rest (TrainStep runs that one beside the stem's forward pass)."""
        if not self.pack_entries:
            return
        head_n, head_work = self.pack_head if self.pack_head else (0, 0)
        n = len(self.pack_entries)
        spans = {None: [(0, head_n, head_work), (head_n, n - head_n, self.n_work - head_work)],
                 0: [(0, head_n, head_work)], 1: [(head_n, n - head_n, self.n_work - head_work)]}[part]
        entry_bytes = ctypes.sizeof(PackEntry)
        for first, count, work in spans:
            if count > 0:
                check(lib().mp_pack_weights(self.flat.data_ptr(), self.packs.data_ptr(),
                                            self.pack_table.data_ptr() + first * entry_bytes, count, work,
                                            stream_ptr(self.device)), 'mp_pack_weights')

    def attach_grads(self):
        """Makes p.grad a view of the flat gradient buffer; zeroes the buffer when the caller has
        cleared the grads (zero_grad(set_to_none=True)), so launches can accumulate (+=)."""
        if all(getattr(s.mod, s.name).grad is None for s in self.params):
            self.flat_grad.zero_()
        for s in self.params:
            p = getattr(s.mod, s.name)
            if p.grad is None:
                if p.requires_grad:
                    p.grad = s.gview
            elif p.grad.data_ptr() != s.gview.data_ptr():
                s.gview.copy_(p.grad)
                p.grad = s.gview


# ---------------------------------------------------------------------------------------- layers
class ConvL:
    def __init__(self, bank, mod, stem=False):
        self.mod, self.stem = mod, stem
        if stem:
            assert mod.kernel_size == (7, 7) and mod.stride == (2, 2) and mod.in_channels == 3
            self.g = C.ConvGeom(147, mod.out_channels, 1)
            master, k = (mod.out_channels, 1, 147), 7
        else:
            tr = isinstance(mod, nn.ConvTranspose2d)
            k = mod.kernel_size[0]
            assert mod.kernel_size == (k, k) and mod.stride[0] == mod.stride[1]
            assert mod.padding == (k // 2, k // 2), 'only "same"-style padding is on the hot path'
            if tr:
                assert mod.stride == (2, 2) and mod.output_padding == (1, 1)
            self.g = C.ConvGeom(mod.in_channels, mod.out_channels, k, mod.stride[0], tr)
            master = self.g.master_shape
        self.w = bank.add_param(mod, 'weight', master_shape=master, conv_k=k)
        self.bias = bank.add_param(mod, 'bias') if mod.bias is not None else None
        g = self.g
        if g.transposed:
            self.fwd = bank.add_matrix(g.cout_p, [(self.w, 1, g.cin, g.cout, g.taps, g.cin_p)])
        else:
            self.fwd = bank.add_matrix(g.cout_p, [(self.w, 0, g.cout, g.cin, g.taps, g.cin_p)])
        self.bwd = None

    def bwd_part(self):
        g = self.g
        if g.transposed:
            return (self.w, 0, g.cin, g.cout, g.taps, g.cout_p)
        return (self.w, 1, g.cout, g.cin, g.taps, g.cout_p)

    def need_bwd(self, bank):
        if self.bwd is None:
            self.bwd = bank.add_matrix(self.g.cin_p, [self.bwd_part()])
        return self.bwd


class BNL:
    def __init__(self, bank, layers, mod, conv_bias=None):
        self.mod = mod
        self.C = mod.num_features
        self.Cp = C.pad64(self.C)
        self.gamma = bank.add_param(mod, 'weight')
        self.beta = bank.add_param(mod, 'bias')
        self.rm = bank.add_buffer(mod, 'running_mean')
        self.rv = bank.add_buffer(mod, 'running_var')
        self.nbt = bank.add_buffer(mod, 'num_batches_tracked')
        self.conv_bias = conv_bias
        self.slot = layers.bn_floats          # fwd sums at slot (2*Cp); saved stats use the same offset
        layers.bn_floats += 2 * self.Cp
        self.index = layers.n_bn              # ticket-counter index
        layers.n_bn += 1
        layers.bns.append(self)


class _NS:
    pass


def _fusable(a, b):
    return a.g.stride == b.g.stride and a.g.transposed == b.g.transposed and \
        a.g.cin_p == b.g.cin_p and a.g.cout_p == b.g.cout_p


def build_layers(model, bank):
    """Layer graph of the reference network (margipose_model.py:103-200) over `bank` storage."""
    L = _NS()
    L.bn_floats = 0
    L.n_bn = 0
    L.bns = []
    inner = model.inner
    L.n_stages, L.n_joints = inner.n_stages, model.n_joints
    cnn = inner.in_cnn
    L.stem = (ConvL(bank, cnn[0], stem=True), BNL(bank, L, cnn[1]))
    L.res_blocks = []
    for layer in (cnn[4], cnn[5]):
        for blk in layer:
            b = _NS()
            names = ['conv1', 'conv2', 'conv3'] if hasattr(blk, 'conv3') else ['conv1', 'conv2']
            b.chain = [(ConvL(bank, getattr(blk, c)), BNL(bank, L, getattr(blk, 'bn' + c[-1]))) for c in names]
            b.down = None
            if blk.downsample is not None:
                b.down = (ConvL(bank, blk.downsample[0]), BNL(bank, L, blk.downsample[1]))
            for conv, _bn in b.chain[1:]:
                conv.need_bwd(bank)
            b.fused = None
            if b.down is not None and _fusable(b.chain[0][0], b.down[0]):
                b.fused = bank.add_matrix(b.chain[0][0].g.cin_p,
                                          [b.chain[0][0].bwd_part(), b.down[0].bwd_part()])
            else:
                b.chain[0][0].need_bwd(bank)
                if b.down is not None:
                    b.down[0].need_bwd(bank)
            L.res_blocks.append(b)
    L.adapter = None
    if len(cnn) > 6:
        conva = ConvL(bank, cnn[6])
        conva.need_bwd(bank)
        L.adapter = (conva, BNL(bank, L, cnn[7], conv_bias=conva.bias))

    def rb_of(block):
        r = _NS()
        r.conv1, r.bn1 = ConvL(bank, block.module[0]), BNL(bank, L, block.module[1])
        r.conv2, r.bn2 = ConvL(bank, block.module[3]), BNL(bank, L, block.module[4])
        r.convs, r.bns = ConvL(bank, block.shortcut[0]), BNL(bank, L, block.shortcut[1])
        r.conv2.need_bwd(bank)
        assert _fusable(r.conv1, r.convs)
        r.fused = bank.add_matrix(r.conv1.g.cin_p, [r.conv1.bwd_part(), r.convs.bwd_part()])
        return r

    bank.mark_pack_head()
    L.columns = []
    L.stage_ranges = []      # [lo, hi) of stage t's parameters in the flat buffers (all-reduce buckets)
    for t in range(L.n_stages):
        row = []
        lo = bank.n_param
        for cols in (inner.xy_hm_cnns, inner.zy_hm_cnns, inner.xz_hm_cnns):
            col = cols[t]
            c = _NS()
            c.mode = {'xy': 0, 'zy': 1, 'xz': 2}[col.heatmap_space]
            c.down = [rb_of(b) for b in col.down_layers]
            c.up = [rb_of(b) for b in col.up_layers]
            c.mid_channels = c.down[-1].conv2.g.cout
            row.append(c)
        L.columns.append(row)
        L.stage_ranges.append((lo, bank.n_param))
    L.combiners = [bank.add_param(cmb.conv, 'weight', master_shape=(cmb.conv.out_channels, 1,
                                                                    cmb.conv.in_channels), conv_k=1)
                   for cmb in inner.hm_combiners]
    return L


class Engine:
    """Buffers + launch programs for one (batch, height, width, mode)."""

    def __init__(self, model, n, h, w, training, device):
        self.model, self.n, self.h, self.w, self.training, self.device = model, n, h, w, training, device
        self.bank, self.L = model._bank, model._layers
        self._bufs = []
        self._keep = []
        # split (bf16x3) precision mode: every activation is a hi / lo pair of bf16 tensors a FIXED distance
        # apart (lo_delta elements), carved from chunks [hi half | lo half]
        self.split = self.bank.split
        self.pairs = self.bank.pairs
        self.lo_delta = 0
        self._chunk = None
        self._chunk_used = 0
        if self.split:
            self.lo_delta = int(os.environ.get('MARGIPOSE_B200_SPLIT_CHUNK', str(1 << 29)))   # elements (1 GiB)
        # one grouped launch per layer for the three columns of a stage (default) or three stream lanes
        self.group = os.environ.get('MARGIPOSE_B200_GROUP', '1') != '0'
        self.streams = [torch.cuda.Stream(device=device) for _ in range(3)]
        # one auxiliary stream per lane: weight gradients (nothing downstream waits for them) and the
        # shortcut conv of a block run beside the lane's dependent chain
        self.aux = [torch.cuda.Stream(device=device) for _ in range(3)]
        # R replicas of the per-BatchNorm fwd sums (2*Cp each, replica stride = bn_floats), followed
        # by the backward reductions (R * 4*Cp per op); producers spread their atomics over replicas
        # (R = 1: the conv epilogue pre-reduces in shared memory and its last CTA finalises the
        # BatchNorm coefficients, so consumers never sum replicas.)  The zeroed-every-step arena also
        # holds the ticket counters: one per BatchNorm (forward) and one per backward reduction.
        # deterministic mode (MargiPoseModel.deterministic, mirrors utils.py:19-24 `init_algorithms(deterministic=True)`):
        # no floating-point atomics between thread blocks anywhere in the step -- batch statistics come from a
        # separate fixed-order pass (mp_bn_stats) into R replicas that mp_bn_fwd adds in order, the backward
        # reductions get one replica per block, weight gradients run without split-K (tunable "deterministic")
        self.det = bool(getattr(model, 'deterministic', False)) or \
            os.environ.get('MARGIPOSE_B200_DETERMINISTIC', '0') == '1'
        # BatchNorm coefficients from the conv epilogue's sums: by the conv's last CTA (True) or by the consuming mp_bn_fwd
        # itself (False: no ticket / fence / finalise tail in the conv kernel)
        self.conv_finalize = (not self.det) and os.environ.get('MARGIPOSE_B200_CONV_FINALIZE', '1') != '0'
        self.R = 32 if self.det else 1     # forward statistics
        self.RB = 32 if self.det else 4    # backward reductions: the blocks spread over RB copies of the sums
        n_stat = (self.R + 2 * self.RB) * self.L.bn_floats
        self.stats = torch.zeros(n_stat + 2 * self.L.n_bn + 8, device=device)
        self._counter_base = n_stat
        self._n_bwd_ops = 0
        self.saves = torch.zeros(max(self.L.bn_floats, 8), device=device)     # mean, 1/std
        self.affine = torch.zeros(max(self.L.bn_floats, 8), device=device)    # scale, shift
        self.coefs = torch.zeros(max(3 * self.L.bn_floats, 8), device=device)  # backward (3, Cp) per branch
        self._stat_n = self.R * self.L.bn_floats
        self._stat_fwd_n = self._stat_n       # [0, _stat_fwd_n): forward sums; [.., _counter_base): backward sums
        self.fwd, self.bwd = [], []
        self.generation = 0    # bumped by every forward(): a backward belongs to exactly one forward
        self.loss_ctx = None   # set by enable_fused_loss(): buffers of the fused loss path
        self._fused = False    # this forward / backward runs the fused loss path (forward(x, fused=True))
        self.bwd_marks = []    # index into self.bwd after which stage t's parameter gradients are complete
        self.trace = []        # (name, buffer, real channels) of every block output, in forward order
        # inference: eval-mode BatchNorm is a per-channel affine, applied in the producing conv's epilogue
        # (bin/infer_single.py:58-66): no BatchNorm pass, no pre-BatchNorm tensor
        self.fold = (not training) and os.environ.get('MARGIPOSE_B200_FOLD', '1') != '0'
        self._u8 = False       # this forward gathers from the uint8 NHWC image buffer (normalisation fused)
        self.x_u8 = None
        self._record()
        if self.fold:
            self._build_fold_table()

    def act(self, n, h, w, c, dtype=torch.bfloat16):
        if not self.split:
            t = torch.zeros(n, h, w, c, dtype=dtype, device=self.device)
            self._bufs.append(t)
            return t
        numel = (n * h * w * c + 7) // 8 * 8
        half = self.lo_delta
        if numel > half:
            raise MargiposeB200Error('an activation of %d elements does not fit a split-mode chunk of %d; raise '
                                     'MARGIPOSE_B200_SPLIT_CHUNK or lower the batch size' % (numel, half))
        if self._chunk is None or self._chunk_used + numel > half:
            self._chunk = torch.zeros(2 * half, dtype=dtype, device=self.device)
            self._chunk_used = 0
            self._bufs.append(self._chunk)
        off = self._chunk_used
        self._chunk_used += numel
        hi = self._chunk[off:off + n * h * w * c].view(n, h, w, c)
        lo = self._chunk[half + off:half + off + n * h * w * c].view(n, h, w, c)
        return self.pairs.register(hi, lo)

    def value(self, t):
        """fp32 value of an activation buffer (hi + lo in split mode)."""
        return self.pairs.value(t) if self.split and t.dtype == torch.bfloat16 else t.float()

    def activation_bytes(self):
        return sum(t.numel() * t.element_size() for t in self._bufs)

    # ---- op recording helpers: an op is a zero-argument callable
    def _launch(self, fn_name, args):
        fn = getattr(lib(), fn_name)
        ref = ctypes.byref(args)
        dev = self.device

        def op():
            rc = fn(ref, stream_ptr(dev))
            if rc != 0:
                check(rc, fn_name)
        op.args = args
        op.name = fn_name
        op.flops = getattr(args, '_flops', 0.0)
        return op

    def conv_ops(self, record):
        """Runs `record()` (calls into convops) capturing its launches as ops."""
        captured = []
        old_i, old_w = C._igemm_launch, C._wgrad_launch
        C._igemm_launch = lambda a, dev: captured.append(self._launch('mp_conv_igemm', a))
        C._wgrad_launch = lambda a, dev: captured.append(self._launch('mp_conv_wgrad', a))
        old_split, C.SPLIT = C.SPLIT, self.pairs
        try:
            record()
        finally:
            C._igemm_launch, C._wgrad_launch = old_i, old_w
            C.SPLIT = old_split
        return captured

    # C-ABI entry points with a grouped variant (<name>_grouped(args[], n, stream))
    GROUPABLE = ('mp_conv_igemm', 'mp_conv_wgrad', 'mp_bn_fwd', 'mp_bn_bwd_reduce', 'mp_bn_bwd_apply', 'mp_bn_stats')

    def _grouped(self, ops):
        """One launch for the same op of the three columns (identical geometry, different tensors)."""
        name = ops[0].name
        fn = getattr(lib(), name + '_grouped')
        arr = (type(ops[0].args) * len(ops))()
        for i, op in enumerate(ops):
            ctypes.memmove(ctypes.addressof(arr[i]), ctypes.addressof(op.args), ctypes.sizeof(op.args))
        n, dev = len(ops), self.device

        def run():
            rc = fn(arr, n, stream_ptr(dev))
            if rc != 0:
                check(rc, name + '_grouped')
        run.name, run.args, run.parts = name, arr, ops     # parts keep the argument structs' referents alive
        run.flops = sum(op.flops for op in ops)
        for flag in ('aux', 'join_aux'):
            if any(getattr(op, flag, False) for op in ops):
                setattr(run, flag, True)
        return run

    def _merge_lanes(self, lanes):
        """The xy / zy / xz columns of a stage (margipose_model.py:196-198) record structurally identical
        programs; zip them into ONE program whose ops are grouped launches (3x the work per launch:
        fixed per-launch costs are paid once, tiles of all three columns fill the SMs together).  Ops
        only some columns have (the axis permutations) keep their own launch."""
        idx, out = [0] * len(lanes), []
        while any(i < len(l) for i, l in zip(idx, lanes)):
            cur = [(k, l[i]) for k, (i, l) in enumerate(zip(idx, lanes)) if i < len(l)]
            names = [getattr(op, 'name', None) for _k, op in cur]
            if len(cur) == len(lanes) and len(set(names)) == 1 and names[0] in self.GROUPABLE:
                out.append(self._grouped([op for _k, op in cur]))
            elif len(cur) == len(lanes) and all(hasattr(op, 'tail') for _k, op in cur) and \
                    len(set(op.tail for _k, op in cur)) == 1:
                kind, stage = cur[0][1].tail
                out.append(self._tail_op(kind, [op.planes for _k, op in cur], stage))
            else:
                solo = [(k, op) for k, op in cur if getattr(op, 'name', None) == 'mp_axis_permute'] or cur
                for k, op in solo:
                    out.append(op)
                cur = solo
            for k, _op in cur:
                idx[k] += 1
        return out

    def _tail_op(self, kind, planes_, stage):
        """Fused-tail launch over up to three planes: ('fwd', logits, prob) / ('bwd', prob, g_in, dlogits).
        With a loss context (enable_fused_loss) and all three planes in the launch this is THE fusion of
        SURVEY.md section 2b (K4 / K5): softmax + expectations + xyz + Gaussian + JS x3 + Euclid, the stage's
        loss accumulated in place; backward = loss gradient + combiner gradient + softmax backward."""
        cols = [list(c) + [None] * (3 - len(planes_)) for c in zip(*planes_)]
        last = stage == self.L.n_stages - 1
        if kind == 'fwd':
            logits, prob = cols

            def op():
                L = self.loss_ctx if self._fused else None
                if L is None or len(planes_) < 3:
                    K._tail_fwd(logits, True, prob=prob)
                else:
                    K._tail_fwd(logits, True, prob=prob, target=L.target, valid_depth=L.valid_depth,
                                coords=L.coords[stage], loss=L.loss_bj, accumulate=stage > 0,
                                pixelwise=L.pixelwise, sigma=L.sigma)
        else:
            prob, g_in, dlogits = cols

            def op():
                L = self.loss_ctx if self._fused else None
                if L is None or len(planes_) < 3:
                    K._tail_bwd(prob, g_in, dlogits, project=True)
                else:   # the last stage has no combiner behind it: no upstream gradient to read
                    K._tail_bwd(prob, None if last else g_in, dlogits, target=L.target, coords=L.coords[stage],
                                w=L.w, valid_depth=L.valid_depth, project=True, pixelwise=L.pixelwise,
                                sigma=L.sigma)
        op.tail = (kind, stage)
        op.planes = tuple(planes_[0])
        return op

    def enable_fused_loss(self, pixelwise=True, sigma=1.0):
        """Static buffers for the fused loss path (train.TrainStep): xyz targets, per-sample valid_depth flags
        (1 = 3D loss, 0 = 2D loss -- the mixed batches of bin/train_3d.py:126-142), joint mask; outputs: the
        per-joint loss summed over stages (margipose_model.py:236-252), per-stage coordinates, the masked
        mean (dsntnn.py:99-121) and its gradient weights."""
        if not self.group:
            raise MargiposeB200Error('the fused loss path needs grouped launches (MARGIPOSE_B200_GROUP=1)')
        n, J, dev = self.n, self.L.n_joints, self.device
        L = _NS()
        L.pixelwise, L.sigma = bool(pixelwise), float(sigma)
        L.target = torch.zeros(n, J, 3, device=dev)
        L.valid_depth = torch.ones(n, dtype=torch.int32, device=dev)
        L.mask = torch.ones(n, J, device=dev)
        L.loss_bj = torch.zeros(n, J, device=dev)
        L.coords = [torch.zeros(n, J, 3, device=dev) for _ in range(self.L.n_stages)]
        L.out2 = torch.zeros(2, device=dev)       # (masked mean, its denominator)
        L.w = torch.zeros(n, J, device=dev)       # d mean / d loss[b, j]
        L.one = torch.ones(1, device=dev)
        self.loss_ctx = L
        return L

    def loss_reduce(self):
        """average_loss over the accumulated per-joint losses and its gradient (two tiny launches)."""
        L, dev = self.loss_ctx, self.device
        nel = L.loss_bj.numel()
        check(lib().mp_masked_mean_fwd(L.loss_bj.data_ptr(), L.mask.data_ptr(), nel, L.out2.data_ptr(),
                                       stream_ptr(dev)), 'mp_masked_mean_fwd')
        if self.training:
            check(lib().mp_masked_mean_bwd(L.one.data_ptr(), L.mask.data_ptr(), L.out2.data_ptr(), nel,
                                           L.w.data_ptr(), stream_ptr(dev)), 'mp_masked_mean_bwd')
        return L.out2[0]

    def _combiner_bwd_op(self, d_next, prev, wc, gin, n, J, hw, cf):
        """HeatmapCombiner backward into the stage's upstream-gradient planes: added to the loss gradients
        the autograd path copied there, or overwriting them on the fused loss path (the tail backward adds
        the loss terms itself)."""
        fn, dev = lib().mp_combiner_bwd, self.device
        pp, gp = planes(prev), planes(gin)

        def op():
            rc = fn(d_next.data_ptr(), pp, wc.data.data_ptr(), gp, wc.grad.data_ptr(),
                    0 if self._fused else 1, n, J, hw, cf, self.lo_delta, stream_ptr(dev))
            if rc != 0:
                check(rc, 'mp_combiner_bwd')
        op.name = 'mp_combiner_bwd'
        op._keep = (pp, gp)
        return op

    def _gather_op(self, patches, n, h, w):
        """ResNet stem conv1 gather from the fp32 NCHW image, or -- forward(uint8 NHWC image) -- from the raw
        image with /255 and ImageNet normalisation fused in (data_specs.py:6-13,38-39)."""
        fn, fn8, dev = lib().mp_stem_im2col, lib().mp_stem_im2col_u8, self.device
        specs = self.model.data_specs.input_specs
        mean = (ctypes.c_float * 3)(*(specs.mean or (0.0, 0.0, 0.0)))
        std = (ctypes.c_float * 3)(*(specs.stddev or (1.0, 1.0, 1.0)))

        def op():
            if self._u8:
                rc = fn8(self.x_u8.data_ptr(), patches.data_ptr(), ctypes.byref(mean), ctypes.byref(std), n, h, w,
                         self.lo_delta, stream_ptr(dev))
            else:
                rc = fn(self.x_in.data_ptr(), patches.data_ptr(), n, h, w, self.lo_delta, stream_ptr(dev))
            if rc != 0:
                check(rc, 'mp_stem_im2col')
        op.name = 'mp_stem_im2col'
        return op

    @staticmethod
    def _on_aux(ops):
        for op in ops:
            op.aux = True
        return ops

    def _moves(self, fn_name, src, dst, *argv):
        """A pure data movement (src, dst, ...): linear, so the split mode runs it once per half of the pair."""
        ops = [self._call(fn_name, src.data_ptr(), dst.data_ptr(), *argv)]
        if self.split:
            ops.append(self._call(fn_name, self.pairs.lo(src).data_ptr(), self.pairs.lo(dst).data_ptr(), *argv))
        return ops

    def _call(self, fn_name, *argv):
        """Op for a C-ABI entry point with scalar/pointer arguments + trailing stream."""
        fn = getattr(lib(), fn_name)
        dev = self.device

        def op():
            rc = fn(*argv, stream_ptr(dev))
            if rc != 0:
                check(rc, fn_name)
        op.name = fn_name
        return op

    def bn_args(self, a, ya, b=None, yb=None, res=None, relu_a=False, relu_out=False, out=None,
                out_nchw=None, hw=0):
        args = BnArgs()
        sbase, vbase = self.stats.data_ptr(), self.saves.data_ptr()

        def fill(br, bn, y):
            br.y = _ptr(y)
            br.sum = sbase + 4 * bn.slot
            br.sq = sbase + 4 * (bn.slot + bn.Cp)
            br.gamma, br.beta = bn.gamma.data.data_ptr(), bn.beta.data.data_ptr()
            br.running_mean, br.running_var = bn.rm.data.data_ptr(), bn.rv.data.data_ptr()
            br.save_mean = vbase + 4 * bn.slot
            br.save_invstd = vbase + 4 * (bn.slot + bn.Cp)
            if self.training and self.conv_finalize:      # finalised by the producing conv's last CTA
                br.scale = self.affine.data_ptr() + 4 * bn.slot
                br.shift = self.affine.data_ptr() + 4 * (bn.slot + bn.Cp)
            if self.training:
                # each BatchNorm is the a- or b-branch of exactly one backward op: 3*Cp coefficients
                br.coef = self.coefs.data_ptr() + 4 * (3 * bn.slot // 2)
            br.conv_bias = bn.conv_bias.data.data_ptr() if bn.conv_bias is not None else None
            br.dgamma, br.dbeta = bn.gamma.grad.data_ptr(), bn.beta.grad.data_ptr()
        fill(args.a, a, ya)
        if b is not None:
            fill(args.b, b, yb)
        args.res = _ptr(res)
        args.relu_a, args.relu_out = int(relu_a), int(relu_out)
        args.out, args.out_nchw = _ptr(out), _ptr(out_nchw)
        args.M = ya.numel() // ya.shape[-1]
        args.C, args.Cp, args.HW = a.C, a.Cp, hw
        args.training = int(self.training)
        args.stat_replicas, args.stat_stride = self.R, self.L.bn_floats
        if a.mod.momentum is None:
            raise MargiposeB200Error('BatchNorm2d(momentum=None) (cumulative moving average) is not on the hot '
                                     'path; the reference uses the default momentum 0.1')
        args.momentum = a.mod.momentum
        args.eps = a.mod.eps
        args.lo_delta = self.lo_delta
        return args

    def stats_of(self, bn):
        s = self.stats
        return (s[bn.slot:bn.slot + bn.Cp], s[bn.slot + bn.Cp:bn.slot + 2 * bn.Cp], self.R, self.L.bn_floats)

    def conv_fwd(self, prog, conv, x, bn):
        g = conv.g
        n, h, w, _ = x.shape
        ho, wo = g.out_hw(h, w)
        y = self.act(n, ho, wo, g.cout_p)
        stats = fin = None
        if self.training and bn is not None and not self.det:
            stats = self.stats_of(bn)
        if stats is not None and self.conv_finalize:
            branch = BnBranch()
            probe = self.bn_args(bn, y)
            ctypes.memmove(ctypes.addressof(branch), ctypes.addressof(probe.a), ctypes.sizeof(BnBranch))
            fin = dict(branch=branch, counter=self.stats.data_ptr() + 4 * (self._counter_base + bn.index),
                       channels=bn.C, count=n * ho * wo, momentum=probe.momentum, eps=probe.eps)
        prog += self.conv_ops(lambda: C.conv_forward(g, x, conv.fwd.t, y, stats=stats, bn=fin))
        if self.training and bn is not None and self.det:    # fixed-order statistics pass over the finished output
            prog.append(self._launch('mp_bn_stats', self.bn_args(bn, y)))
        return y

    def bn_bwd(self, prog, fwd_args, dout=None, dout_nchw=None, dya=None, dyb=None, dres=None):
        """Backward of a bn_fwd op: same argument struct plus the gradient pointers."""
        args = BnArgs.from_buffer_copy(fwd_args)
        args.dout, args.dout_nchw = _ptr(dout), _ptr(dout_nchw)
        args.a.dy, args.b.dy, args.dres = _ptr(dya), _ptr(dyb), _ptr(dres)
        args.stat_replicas = self.RB
        args.sums = self.stats.data_ptr() + 4 * self._stat_n
        # (bwd_counter stays NULL: finalising the coefficients in the reduce pass's last block costs
        # more -- every block must fence its atomics -- than the one-phase prologue of the apply pass)
        self._stat_n += 4 * fwd_args.Cp * self.RB
        assert self._stat_n <= self._counter_base
        prog.append(self._launch('mp_bn_bwd_reduce', args))
        prog.append(self._launch('mp_bn_bwd_apply', args))

    # ---- inference with folded BatchNorm
    def _build_fold_table(self):
        bns = self.L.bns
        table = (BnFoldEntry * len(bns))()
        base = self.affine.data_ptr()
        for e, bn in zip(table, bns):
            e.gamma, e.beta = bn.gamma.data.data_ptr(), bn.beta.data.data_ptr()
            e.running_mean, e.running_var = bn.rm.data.data_ptr(), bn.rv.data.data_ptr()
            e.conv_bias = bn.conv_bias.data.data_ptr() if bn.conv_bias is not None else None
            e.scale, e.shift = base + 4 * bn.slot, base + 4 * (bn.slot + bn.Cp)
            e.C, e.Cp, e.eps = bn.C, bn.Cp, bn.mod.eps
        raw = torch.frombuffer(bytearray(bytes(table)), dtype=torch.uint8)
        self._fold_table = raw.to(self.device)
        self._n_fold = len(bns)

    def conv_folded(self, prog, conv, x, bn, relu, res=None):
        """y = relu?(BatchNorm_eval(conv(x))) (+ res): one launch, the affine rides in the conv epilogue."""
        g = conv.g
        n, h, w, _ = x.shape
        ho, wo = g.out_hw(h, w)
        y = self.act(n, ho, wo, g.cout_p)
        base = self.affine.data_ptr()
        ep = (base + 4 * bn.slot, base + 4 * (bn.slot + bn.Cp), relu)
        prog += self.conv_ops(lambda: C.conv_forward(g, x, conv.fwd.t, y, res=res, ep=ep))
        return y

    @staticmethod
    def _foldable(conv):
        # a 1x1 transposed conv writes one of four output parity classes; BatchNorm's shift belongs on all four
        return not (conv.g.transposed and conv.g.k == 1)

    # ---- blocks: each returns (output buffer, backward builder)
    def residual_block(self, fwd, rb, x, logits_out=None):
        """MargiPose ResidualBlock (margipose_model.py:25-40)."""
        if self.fold and logits_out is None and self._foldable(rb.convs):
            ys = self.conv_folded(fwd, rb.convs, x, rb.bns, 0)                 # bn_s(conv_s(x))
            a1 = self.conv_folded(fwd, rb.conv1, x, rb.bn1, 1)                 # relu(bn_1(conv_1(x)))
            return self.conv_folded(fwd, rb.conv2, a1, rb.bn2, 1, res=ys), None   # relu(bn_2(conv_2(a1))) + shortcut
        side = []
        ys = self.conv_fwd(side, rb.convs, x, rb.bns)
        fwd += self._on_aux(side)
        y1 = self.conv_fwd(fwd, rb.conv1, x, rb.bn1)
        a1 = self.act(*y1.shape)
        f1 = self.bn_args(rb.bn1, y1, relu_a=True, out=a1)
        fwd.append(self._launch('mp_bn_fwd', f1))
        y2 = self.conv_fwd(fwd, rb.conv2, a1, rb.bn2)
        n, h, w, cp = y2.shape
        out = None if logits_out is not None else self.act(n, h, w, cp)
        f2 = self.bn_args(rb.bn2, y2, b=rb.bns, yb=ys, relu_a=True, out=out, out_nchw=logits_out, hw=h * w)
        fwd.append(self._launch('mp_bn_fwd', f2))
        fwd[-1].join_aux = True

        def backward(bwd, dout=None, dout_nchw=None, need_dx=True):
            dy2, dys = self.act(*y2.shape), self.act(*ys.shape)
            self.bn_bwd(bwd, f2, dout=dout, dout_nchw=dout_nchw, dya=dy2, dyb=dys)
            bwd += self._on_aux(self.conv_ops(lambda: C.conv_wgrad(rb.conv2.g, a1, dy2, rb.conv2.w.grad)))
            bwd += self._on_aux(self.conv_ops(lambda: C.conv_wgrad(rb.convs.g, x, dys, rb.convs.w.grad)))
            da1 = self.act(*a1.shape)
            bwd += self.conv_ops(lambda: C.conv_dgrad(rb.conv2.g, dy2, rb.conv2.bwd.t, da1))
            dy1 = self.act(*y1.shape)
            self.bn_bwd(bwd, f1, dout=da1, dya=dy1)
            bwd += self._on_aux(self.conv_ops(lambda: C.conv_wgrad(rb.conv1.g, x, dy1, rb.conv1.w.grad)))
            if not need_dx:
                return None
            dx = self.act(*x.shape)
            bwd += self.conv_ops(lambda: C.conv_dgrad(rb.conv1.g, dy1, rb.fused.t, dx,
                                                      second=(rb.convs.g, dys, rb.fused.koffs[1])))
            return dx
        return (logits_out if logits_out is not None else out), backward

    def resnet_block(self, fwd, blk, x):
        """torchvision BasicBlock / Bottleneck: relu(bn_last(conv_last(...)) + identity)."""
        chain, down = blk.chain, blk.down
        if self.fold:
            ident = x if down is None else self.conv_folded(fwd, down[0], x, down[1], 0)
            a = x
            for conv, bn in chain[:-1]:
                a = self.conv_folded(fwd, conv, a, bn, 1)
            return self.conv_folded(fwd, chain[-1][0], a, chain[-1][1], 2, res=ident), None   # relu(bn(conv) + identity)
        acts, ys, fargs = [x], [], []
        for i, (conv, bn) in enumerate(chain):
            y = self.conv_fwd(fwd, conv, acts[-1], bn)
            ys.append(y)
            if i < len(chain) - 1:
                a = self.act(*y.shape)
                f = self.bn_args(bn, y, relu_a=True, out=a)
                fwd.append(self._launch('mp_bn_fwd', f))
                fargs.append(f)
                acts.append(a)
        out = self.act(*ys[-1].shape)
        if down is not None:
            yd = self.conv_fwd(fwd, down[0], x, down[1])
            fl = self.bn_args(chain[-1][1], ys[-1], b=down[1], yb=yd, relu_out=True, out=out)
        else:
            yd = None
            fl = self.bn_args(chain[-1][1], ys[-1], res=x, relu_out=True, out=out)
        fwd.append(self._launch('mp_bn_fwd', fl))

        def backward(bwd, dout, need_dx=True):
            dyl = self.act(*ys[-1].shape)
            dyd = self.act(*yd.shape) if yd is not None else None
            dz = self.act(*x.shape) if (yd is None and need_dx) else None
            self.bn_bwd(bwd, fl, dout=dout, dya=dyl, dyb=dyd, dres=dz)
            dy = dyl
            for i in range(len(chain) - 1, 0, -1):
                conv = chain[i][0]
                bwd += self._on_aux(self.conv_ops(lambda conv=conv, i=i, dy=dy: C.conv_wgrad(conv.g, acts[i], dy, conv.w.grad)))
                da = self.act(*acts[i].shape)
                bwd += self.conv_ops(lambda conv=conv, dy=dy, da=da: C.conv_dgrad(conv.g, dy, conv.bwd.t, da))
                dy = self.act(*ys[i - 1].shape)
                self.bn_bwd(bwd, fargs[i - 1], dout=da, dya=dy)
            conv0 = chain[0][0]
            bwd += self._on_aux(self.conv_ops(lambda: C.conv_wgrad(conv0.g, x, dy, conv0.w.grad)))
            if down is not None:
                bwd += self._on_aux(self.conv_ops(lambda: C.conv_wgrad(down[0].g, x, dyd, down[0].w.grad)))
            if not need_dx:
                return None
            dx = self.act(*x.shape)
            if down is None:
                bwd += self.conv_ops(lambda: C.conv_dgrad(conv0.g, dy, conv0.bwd.t, dx, res=dz))
            elif blk.fused is not None:
                bwd += self.conv_ops(lambda: C.conv_dgrad(conv0.g, dy, blk.fused.t, dx,
                                                          second=(down[0].g, dyd, blk.fused.koffs[1])))
            else:   # different strides: full dgrad first, then the strided one accumulates in place
                bwd += self.conv_ops(lambda: C.conv_dgrad(conv0.g, dy, conv0.bwd.t, dx))
                bwd += self.conv_ops(lambda: C.conv_dgrad(down[0].g, dyd, down[0].bwd.t, dx, res=dx))
            return dx
        return out, backward

    # ---- whole network
    def _record(self):
        L, dev, n, h, w = self.L, self.device, self.n, self.h, self.w
        training = self.training
        J = L.n_joints
        fwd = []
        # ---- stem (margipose_model.py:119-138)
        self.x_in = torch.zeros(n, 3, h, w, device=dev)
        patches = self.act(n, h // 2, w // 2, 192)
        fwd.append(self._gather_op(patches, n, h, w))
        conv0, bn0 = L.stem
        if self.fold:
            y0 = f0 = None
            a0 = self.conv_folded(fwd, conv0, patches, bn0, 1)
        else:
            y0 = self.conv_fwd(fwd, conv0, patches, bn0)
            a0 = self.act(*y0.shape)
            f0 = self.bn_args(bn0, y0, relu_a=True, out=a0)
            fwd.append(self._launch('mp_bn_fwd', f0))
        hp, wp, c0 = h // 4, w // 4, a0.shape[-1]
        p0 = self.act(n, hp, wp, c0)
        idx = torch.zeros(n, hp, wp, c0, dtype=torch.uint8, device=dev)
        self._bufs.append(idx)
        fwd.append(self._call('mp_maxpool_fwd', a0.data_ptr(), p0.data_ptr(), idx.data_ptr(), n, h // 2, w // 2, c0,
                              self.lo_delta))
        self.trace += [('stem.relu', a0, bn0.C), ('stem.maxpool', p0, bn0.C)]
        x = p0
        res_bwd = []
        for i, blk in enumerate(L.res_blocks):
            x, b = self.resnet_block(fwd, blk, x)
            res_bwd.append(b)
            self.trace.append(('resnet.%d' % i, x, blk.chain[-1][1].C))
        x_pre = x
        if L.adapter is not None:
            conva, bna = L.adapter
            if self.fold:
                x = self.conv_folded(fwd, conva, x, bna, 1)
            else:
                ya = self.conv_fwd(fwd, conva, x, bna)
                x = self.act(*ya.shape)
                fa = self.bn_args(bna, ya, relu_a=True, out=x)
                fwd.append(self._launch('mp_bn_fwd', fa))
            self.trace.append(('adapter', x, bna.C))
        feats = x
        hf, wf, cf = feats.shape[1], feats.shape[2], feats.shape[3]
        self.heatmap_hw = (hf, wf)
        # ---- stages (margipose_model.py:188-199)
        segs = [('serial', fwd)]
        inp = feats
        self.probs, self.logits = [], []
        inps, col_bwd, comb = [], [], []
        for t in range(L.n_stages):
            if t > 0:
                prev, wc = self.probs[t - 1], L.combiners[t - 1]
                new_inp = self.act(*inp.shape)
                segs.append(('serial', [self._call('mp_combiner_fwd', planes(prev), wc.data.data_ptr(),
                                                   inp.data_ptr(), new_inp.data_ptr(), n, J, hf * wf, cf,
                                                   self.lo_delta)]))
                comb.append((prev, wc))
                inp = new_inp
            inps.append(inp)
            lanes, lane_bwd, lz, pr = [], [], [], []
            for col in L.columns[t]:
                ops, blk_b, perm = [], [], None
                xcol = inp
                for i, rb in enumerate(col.down):
                    xcol, b = self.residual_block(ops, rb, xcol)
                    blk_b.append(b)
                    self.trace.append(('stage%d.col%d.down%d' % (t, col.mode, i), xcol, rb.bn2.C))
                if col.mode != 0:
                    s, cmid = xcol.shape[1], col.mid_channels
                    if xcol.shape[1] != xcol.shape[2] or cmid % s != 0:
                        raise MargiposeB200Error(
                            'axis permutation needs a square mid feature map whose side (%d) divides '
                            'the channel count (%d)' % (s, cmid))
                    xp = self.act(*xcol.shape)
                    ops += self._moves('mp_axis_permute', xcol, xp, col.mode, n, s, cmid, xcol.shape[-1])
                    perm = (col.mode, s, cmid, tuple(xcol.shape))
                    xcol = xp
                logits = torch.zeros(n, J, hf, wf, device=dev)
                for i, rb in enumerate(col.up):
                    last = i == len(col.up) - 1
                    xcol, b = self.residual_block(ops, rb, xcol, logits_out=logits if last else None)
                    blk_b.append(b)
                    self.trace.append(('stage%d.col%d.up%d' % (t, col.mode, i), xcol, rb.bn2.C))
                prob = torch.zeros(n, J, hf, wf, device=dev)
                ops.append(self._tail_op('fwd', [(logits, prob)], t))
                lanes.append(ops)
                lz.append(logits)
                pr.append(prob)
                lane_bwd.append((blk_b, perm, len(col.down)))
            segs.append(('serial', self._merge_lanes(lanes)) if self.group else ('parallel', lanes))
            self.probs.append(pr)
            self.logits.append(lz)
            col_bwd.append(lane_bwd)
        self.fwd = segs
        self.stage_inputs = inps     # feats + sum of the earlier stages' combiner outputs (layer-wise parity tests)
        if not training:
            return

        # ---- backward program
        bsegs = []
        self.gin = [[torch.zeros(n, J, hf, wf, device=dev) for _ in range(3)] for _ in range(L.n_stages)]
        d_next = None
        for t in range(L.n_stages - 1, -1, -1):
            if t < L.n_stages - 1:
                prev, wc = comb[t]
                bsegs.append(('serial', [self._combiner_bwd_op(d_next, prev, wc, self.gin[t], n, J, hf * wf, cf)]))
            lanes, dxs = [], []
            for k in range(3):
                ops = []
                blk_b, perm, n_down = col_bwd[t][k]
                prob, g_in = self.probs[t][k], self.gin[t][k]
                dlogits = torch.zeros(n, J, hf, wf, device=dev)
                ops.append(self._tail_op('bwd', [(prob, g_in, dlogits)], t))
                d = None
                for i in range(len(blk_b) - 1, -1, -1):
                    if i == len(blk_b) - 1:
                        d = blk_b[i](ops, dout_nchw=dlogits)
                    else:
                        d = blk_b[i](ops, dout=d)
                    if perm is not None and i == n_down:
                        mode, s, cmid, shp = perm
                        dp = self.act(*shp)
                        ops += self._moves('mp_axis_permute', d, dp, mode, n, s, cmid, shp[-1])
                        d = dp
                lanes.append(ops)
                dxs.append(d)
            bsegs.append(('serial', self._merge_lanes(lanes)) if self.group else ('parallel', lanes))
            d_inp = self.act(*inps[t].shape)
            terms = dxs + ([d_next] if d_next is not None else [])
            arr = (ctypes.c_void_p * 4)(*([x.data_ptr() for x in terms] + [None] * (4 - len(terms))))
            bsegs.append(('serial', [self._call('mp_add_bf16', ctypes.byref(arr), len(terms), d_inp.data_ptr(),
                                                d_inp.numel(), self.lo_delta)]))
            self._keep.append(arr)
            d_next = d_inp
            self.bwd_marks.append(len(bsegs))    # stage t's (and combiner t's) parameter gradients are complete
        # ---- stem backward
        ops = []
        d = d_next
        if L.adapter is not None:
            conva, bna = L.adapter
            dya = self.act(*ya.shape)
            self.bn_bwd(ops, fa, dout=d, dya=dya)
            ops += self.conv_ops(lambda: C.conv_wgrad(conva.g, x_pre, dya, conva.w.grad))
            d = self.act(*x_pre.shape)
            ops += self.conv_ops(lambda: C.conv_dgrad(conva.g, dya, conva.bwd.t, d))
        for i in range(len(res_bwd) - 1, -1, -1):
            d = res_bwd[i](ops, d, need_dx=True)
        da0 = self.act(*a0.shape)
        ops.append(self._call('mp_maxpool_bwd', d.data_ptr(), idx.data_ptr(), da0.data_ptr(), n, h // 2, w // 2, c0))
        if self.split:
            ops.append(self._call('mp_maxpool_bwd', self.pairs.lo(d).data_ptr(), idx.data_ptr(),
                                  self.pairs.lo(da0).data_ptr(), n, h // 2, w // 2, c0))
        dy0 = self.act(*y0.shape)
        self.bn_bwd(ops, f0, dout=da0, dya=dy0)
        ops += self.conv_ops(lambda: C.conv_wgrad(conv0.g, patches, dy0, conv0.w.grad))
        bsegs.append(('serial', ops))
        self.bwd = bsegs

    # ---- execution
    def _run_lane(self, ops, main, aux):
        """Ops in program order on `main` (the current stream); ops flagged .aux go to `aux` after
        waiting for everything issued on `main` so far; an op flagged .join_aux first waits for aux."""
        pending = False
        for op in ops:
            if getattr(op, 'aux', False):
                aux.wait_stream(main)
                with torch.cuda.stream(aux):
                    op()
                pending = True
            else:
                if pending and getattr(op, 'join_aux', False):
                    main.wait_stream(aux)
                    pending = False
                op()
        if pending:
            main.wait_stream(aux)

    def _run(self, segs, use_aux=True):
        cur = torch.cuda.current_stream(self.device)
        for kind, body in segs:
            if kind == 'serial':
                self._run_lane(body, cur, self.aux[0] if use_aux else cur)
            else:
                for s, a, ops in zip(self.streams, self.aux, body):
                    s.wait_stream(cur)
                    with torch.cuda.stream(s):
                        self._run_lane(ops, s, a if use_aux else s)
                for s in self.streams:
                    cur.wait_stream(s)

    def time_kernel_class(self, name, reps=3):
        """GPU time of every launch of C-ABI entry point `name` in one step (forward + backward
        programs), run back to back in program order on one stream and replayed from a CUDA graph --
        i.e. each kernel alone on the GPU, no host overhead.  Returns (launches, total ms per step,
        total algorithmic flops per step)."""
        ops = []
        for segs in (self.fwd, self.bwd):
            for kind, body in segs:
                for lane in ([body] if kind == 'serial' else body):
                    ops += [op for op in lane if getattr(op, 'name', 'tail') == name]
        if not ops:
            return 0, 0.0, 0.0
        for op in ops:
            op()
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for op in ops:
                op()
        g.replay()
        torch.cuda.synchronize(self.device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize(self.device)
        return len(ops), e0.elapsed_time(e1) / reps, sum(getattr(op, 'flops', 0.0) for op in ops)

    def launches(self, segs=None):
        out = 0
        for prog in ([self.fwd, self.bwd] if segs is None else [segs]):
            for kind, body in prog:
                out += len(body) if kind == 'serial' else sum(len(o) for o in body)
        return out

    def forward(self, x, fused=False, join_after_stem=None):
        """x: fp32 (N, 3, H, W) on the device.  Returns probs[t][k], fp32 (N, J, h, w).  fused=True (after
        enable_fused_loss): the stage tails also compute the losses of the loss context.  The returned
        join_after_stem: a stream the pass waits for between the feature extractor and the first stage.  The returned
        tensors (and everything a backward reads) are the engine's static buffers: the next forward of this
        engine overwrites them, and a backward is only valid for the most recent forward (checked through
        `generation`)."""
        if x.dtype == torch.uint8:      # raw NHWC image: the stem gather normalises on the fly
            if tuple(x.shape) != (self.n, self.h, self.w, 3):
                raise MargiposeB200Error('uint8 input must be (N, H, W, 3) = %s, got %s'
                                         % ((self.n, self.h, self.w, 3), tuple(x.shape)))
            if self.x_u8 is None:
                self.x_u8 = torch.zeros(self.n, self.h, self.w, 3, dtype=torch.uint8, device=self.device)
            if x.data_ptr() != self.x_u8.data_ptr():
                self.x_u8.copy_(x)
            self._u8 = True
        else:
            if x.data_ptr() != self.x_in.data_ptr():
                self.x_in.copy_(x)
            self._u8 = False
        self.generation += 1
        lib().mp_set_tunable(b'deterministic', int(self.det))
        self._fused = bool(fused)
        if fused and self.loss_ctx is None:
            raise MargiposeB200Error('forward(fused=True) needs enable_fused_loss() first')
        if self.training:
            self.stats[:self._stat_fwd_n].zero_()       # forward BatchNorm sums
            self.stats[self._counter_base:].zero_()     # ticket counters
            self.bank.flat_cnt.add_(1)
        if self.fold:   # running statistics -> per-channel affines for every BatchNorm of the network, one launch
            check(lib().mp_bn_fold_eval(self._fold_table.data_ptr(), self._n_fold, stream_ptr(self.device)),
                  'mp_bn_fold_eval')
        if join_after_stem is None:
            self._run(self.fwd)
        else:   # the columns' weight packs are being refreshed on another stream while the feature extractor runs
            self._run(self.fwd[:1])
            torch.cuda.current_stream(self.device).wait_stream(join_after_stem)
            self._run(self.fwd[1:])
        return self.probs

    def backward(self, grads=None, lo=0, hi=None):
        """grads[t][k]: fp32 (N, J, h, w) gradient w.r.t. the stage-t plane-k heatmap, or None; grads=None on
        the fused loss path (the tail backward derives the loss gradient itself).  [lo, hi) selects a slice of
        the backward program (see bwd_marks) so a caller can start all-reducing finished parameter ranges."""
        lib().mp_set_tunable(b'deterministic', int(self.det))
        if lo == 0:
            # backward reductions are atomically accumulated: zero them per backward (a second backward over
            # the same forward -- retain_graph, separate 2D / 3D loss backwards -- must not see the first's sums)
            self.stats[self._stat_fwd_n:self._counter_base].zero_()
            if not self._fused:
                for t, row in enumerate(self.gin):
                    for k, g in enumerate(row):
                        if grads is None or grads[t][k] is None:
                            g.zero_()
                        else:
                            g.copy_(grads[t][k])
        self._run(self.bwd[lo:hi])
