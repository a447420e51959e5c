"""model/updater.py of raahii/mocogan-chainer, re-hosted on libmcg.so (updater.py:9-113).

Same constructor kwargs and method names.  One forward, then the three loss -> cleargrads -> backward -> WeightDecay
-> Adam passes in the reference's order A (image_dis), B (video_dis), C (image_gen), with in-place weight updates and
no re-run of the discriminators, so pass C differentiates through updated weights with the activations of the
forward (SURVEY.md §3.2 pt 3).  Back-propagation whose result the reference discards is not executed (pts 2, 4).
"""
import numpy as np
import torch

from .. import chainer
from .. import random as mrandom
from ..chainer import Variable
from ..chainer import functions as F
from ..chainer.dataset import concat_examples


class Updater(chainer.training.StandardUpdater):
    def __init__(self, *args, **kwargs):
        self.model = kwargs.pop('model')
        self.image_gen, self.image_dis, self.video_dis = kwargs.pop('models')
        self.video_length = kwargs.pop('video_length')
        self.img_size = kwargs.pop('img_size')
        self.channel = kwargs.pop('channel')
        self.dim_zl = kwargs.pop('dim_zl')
        self.tf_writer = kwargs.pop('tensorboard_writer')
        # additive (not in the reference): CUDA-graph capture of the device part of the step
        self.use_graph = kwargs.pop('use_graph', False)
        self.graph_warmup = kwargs.pop('graph_warmup', 2)
        # additive: run the (small) image discriminator on a second stream, concurrently with the video discriminator
        self.use_streams = kwargs.pop('use_streams', True)
        self._side = None
        # additive: in graph mode, copy batch i+1 host->device on a copy stream while step i runs
        self.prefetch = kwargs.pop('prefetch', True)
        self._stage, self._stage_next, self._copy_stream, self._flags = None, None, None, None
        # additive: parity tests set this to read the forward's Variables (x_fake, y_*) after a step
        self.keep_forward = kwargs.pop('keep_forward', False)
        self.last_forward = None

        super(Updater, self).__init__(*args, **kwargs)
        self.losses = {}
        self._graph = None
        self._static = None
        self._eager_steps = 0

    # ------------------------------------------------------------------ losses
    def loss_dis(self, dis, y_real, y_fake, t_real, t_fake):
        # gan criterion + (infogan, VideoDiscriminator only) categorical criterion — updater.py:25-37, one kernel
        use_ce = self.model == 'infogan' and dis.name == "VideoDiscriminator"
        loss = F.gan_loss_dis(y_real, y_fake, _lab(t_real), _lab(t_fake), use_ce)
        self.losses[dis.name] = loss.data
        self._report(dis, loss)
        return loss

    def loss_gen(self, gen, y_fake_i, y_fake_v, t_fake):
        # updater.py:50-56, one kernel
        loss = F.gan_loss_gen(y_fake_i, y_fake_v, _lab(t_fake), self.model == 'infogan')
        self.losses[gen.name] = loss.data
        self._report(gen, loss)
        return loss

    def _report(self, link, loss):
        # updater.py:39-42,58-61: only on a new epoch, so the loss scalar is read back once per epoch, not per step
        if self.is_new_epoch and self._graph is None and not torch.cuda.is_current_stream_capturing():
            chainer.report({'loss': loss.data}, link)
            if self.tf_writer is not None:
                self.tf_writer.add_scalar('loss:{}'.format(link.name), float(loss.data), self.epoch)

    def _report_replayed(self):
        """Graph mode: the loss nodes are not re-executed on replay, so the new-epoch report of updater.py:39-42,58-61
        reads the captured step's static loss tensors instead."""
        if not self.is_new_epoch:
            return
        for link in (self.image_dis, self.video_dis, self.image_gen):
            loss = self.losses.get(link.name)
            if loss is None:
                continue
            chainer.report({'loss': loss}, link)
            if self.tf_writer is not None:
                self.tf_writer.add_scalar('loss:{}'.format(link.name), float(loss), self.epoch)

    def concat_label_video(self, video, label, xp=None):
        """updater.py:65-76 (cgan): append dim_zl planes of -1 with +1 at the label plane.  A FunctionNode, so the fake
        clip stays attached to the generator (F.concat in the reference)."""
        return F.concat_label_video(video, label, self.dim_zl)

    # ------------------------------------------------------------------ one step on device-resident inputs
    def step_on_device(self, x_real, t_real):
        """updater.py:96-113 given the real batch already on the device: x_real (N,C,T,H,W), t_real int32 (N) | None."""
        image_gen_optimizer = self.get_optimizer('image_gen')
        image_dis_optimizer = self.get_optimizer('image_dis')
        video_dis_optimizer = self.get_optimizer('video_dis')
        image_gen = self.image_gen
        image_dis, video_dis = self.image_dis, self.video_dis
        src = mrandom.get_source()
        src.begin_step()
        batchsize = x_real.shape[0]

        x_real = Variable(x_real, requires_grad=False)   # the reference's default requires_grad=True is dead work
        t_real = None if t_real is None else Variable(t_real, requires_grad=False)
        if self.model == 'cgan':
            x_real = self.concat_label_video(x_real, t_real)
        t = src.frame()                                   # updater.py:96, stays on the device
        # Streams (additive): the step's independent branches run side by side — image discriminator on `di`, the
        # generator on `g`, the video discriminator's fake-clip branch on `dvf`, its real-clip branch (and everything
        # else) on the caller's stream.  Every FunctionNode remembers its stream, and Variable.backward() replays the
        # node there (chainer.config.branch_streams), so e.g. pass B's real and fake chains interleave convolution
        # kernels (tensor-bound) with BatchNorm / activation passes (HBM-bound) of the other chain.
        main = torch.cuda.current_stream()
        if self.use_streams:
            if self._side is None:
                self._side = {k: torch.cuda.Stream() for k in ("di", "g", "dvf")}
            st_di, st_g, st_dvf = self._side["di"], self._side["g"], self._side["dvf"]
        else:
            st_di = st_g = st_dvf = main
        old_branch = chainer.config.branch_streams
        chainer.config.branch_streams = bool(self.use_streams)
        try:
            for s_ in {st_di, st_g, st_dvf} - {main}:
                s_.wait_stream(main)
            with torch.cuda.stream(st_di):
                y_real_i = image_dis(x_real, frame=t)     # updater.py:97
            y_real_v = video_dis(x_real)                  # updater.py:98
            dv_real_done = torch.cuda.Event()
            dv_real_done.record(main)

            ## fake data
            with torch.cuda.stream(st_g):
                x_fake, t_fake = image_gen(batchsize)         # updater.py:101  (T,N,C,H,W)
                x_fake = x_fake.transpose(1, 2, 0, 3, 4)      # updater.py:102  (N,C,T,H,W), still attached to G
            t_fake = None if t_fake is None else Variable(t_fake, requires_grad=False)
            x_gen = x_fake                                # what passes A/B must not back-propagate past
            if self.model == 'cgan':
                with torch.cuda.stream(st_g):
                    x_fake = self.concat_label_video(x_fake, t_fake)   # updater.py:104-106
                x_gen = x_fake
            st_di.wait_stream(st_g)
            with torch.cuda.stream(st_di):
                y_fake_i = image_dis(x_fake, frame=t)     # updater.py:107
            st_dvf.wait_stream(st_g)
            st_dvf.wait_event(dv_real_done)               # BatchNorm running statistics: real call first, then fake
            with torch.cuda.stream(st_dvf):
                y_fake_v = video_dis(x_fake)              # updater.py:108

            if self.keep_forward:
                self.last_forward = dict(x_fake=x_gen, y_real_i=y_real_i, y_real_v=y_real_v, y_fake_i=y_fake_i,
                                         y_fake_v=y_fake_v)

            ## update  (updater.py:111-113)
            # passes A, B: gradients flowing from the discriminator losses into the generator are discarded by the
            # reference (image_gen.cleargrads() in pass C) -> not computed.  pass C: discriminator wgrads are discarded.
            # Passes A and B touch disjoint parameters and activations, so A runs on `di` while B runs on the caller's
            # stream (+ `dvf`); both are complete before pass C reads the updated discriminator weights.
            image_dis_optimizer.stop_variables = video_dis_optimizer.stop_variables = (x_gen,)
            image_gen_optimizer.frozen_links = (image_dis, video_dis)
            with torch.cuda.stream(st_di):
                image_dis_optimizer.update(self.loss_dis, image_dis, y_real_i, y_fake_i, t_real, t_fake)
            main.wait_stream(st_dvf)
            # the video discriminator's all-reduce + Adam go on `dvf`: in pass C only the Dv branch (which runs there)
            # needs the updated weights, so the image-discriminator branch of pass C starts under them
            video_dis_optimizer.update_stream = st_dvf if self.use_streams else None
            video_dis_optimizer.update(self.loss_dis, video_dis, y_real_v, y_fake_v, t_real, t_fake)
            main.wait_stream(st_di)
            main.wait_stream(st_g)
            image_gen_optimizer.update(self.loss_gen, image_gen, y_fake_i, y_fake_v, t_fake)
            for s_ in {st_di, st_g, st_dvf} - {main}:
                main.wait_stream(s_)
        finally:
            chainer.config.branch_streams = old_branch

    # ------------------------------------------------------------------ update_core
    def _next_host_batch(self):
        it = self.get_iterator('main')
        batch = it.next()
        x_real, t_real = concat_examples(batch)
        if t_real is not None and not torch.is_tensor(t_real):
            t_real = np.asarray(t_real).astype(np.int32)
        # a batch that lives in re-used pinned staging memory (datasets.Uint8ClipCache) wants to know when its
        # host->device copy has executed, so that the buffer is not refilled under a copy still queued on the device
        self._copied_cb = getattr(batch, "copied", None)
        return x_real, t_real, (it.is_new_epoch, it.epoch)

    def _report_copied(self, stream=None):
        cb, self._copied_cb = getattr(self, "_copied_cb", None), None
        if cb is not None:
            ev = torch.cuda.Event()
            ev.record(stream if stream is not None else torch.cuda.current_stream())
            cb(ev)

    @property
    def is_new_epoch(self):
        return self._flags[0] if self._flags is not None else self._iterators['main'].is_new_epoch

    @property
    def epoch(self):
        return self._flags[1] if self._flags is not None else self._iterators['main'].epoch

    def update_core(self):
        ## real data  (updater.py:87-92)
        if self.use_graph and self.prefetch:
            return self._update_core_prefetched()
        x_real, t_real, _ = self._next_host_batch()
        if self.use_graph:
            return self.step_host_inputs(x_real, t_real)
        x_real = self.converter(x_real, self.device)
        t_real = None if t_real is None else self.converter(t_real, self.device)
        self._report_copied()
        self.step_on_device(x_real, t_real)

    def _issue_h2d(self, k, host):
        """Starts the host->device copy of one batch into staging slot k on the copy stream."""
        as_t = lambda a: a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a))
        x, t, flags = host
        x = as_t(x)
        t = None if t is None else as_t(t)
        slot = self._stage[k]
        if slot is None:
            dev = torch.device("cuda", torch.cuda.current_device())
            # same strides as the host batch (the uint8 clip cache hands out a channels-last view): a plain memcpy
            slot = self._stage[k] = {"x": torch.empty_strided(x.shape, x.stride(), dtype=x.dtype, device=dev),
                                     "t": None if t is None else torch.empty(t.shape, dtype=torch.int32, device=dev),
                                     "ready": torch.cuda.Event(), "consumed": None}
        cs = self._copy_stream
        if slot["consumed"] is not None:
            cs.wait_event(slot["consumed"])        # the step that read this slot has copied it out
        with torch.cuda.stream(cs):
            slot["x"].copy_(x, non_blocking=True)
            if t is not None:
                slot["t"].copy_(t, non_blocking=True)
            slot["ready"].record(cs)
        self._report_copied(cs)
        slot["flags"] = flags

    def _update_core_prefetched(self):
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream()
            self._stage = [None, None]
            self._stage_next = 0
            self._issue_h2d(0, self._next_host_batch())
        k = self._stage_next
        slot = self._stage[k]
        main = torch.cuda.current_stream()
        main.wait_event(slot["ready"])
        self._flags = slot["flags"]
        self._stage_next = 1 - k
        self.step_host_inputs(slot["x"], slot["t"])           # device->static copy + graph replay on `main` (asynchronous)
        ev = torch.cuda.Event()
        ev.record(main)
        slot["consumed"] = ev
        # the host assembles batch i+1 (sub-sequence draws, gather into pinned memory) and starts its H2D copy while the
        # device runs step i
        self._issue_h2d(1 - k, self._next_host_batch())

    def step_host_inputs(self, x_real, t_real):
        """CUDA-graph path: the batch (host or device) is copied straight into static device buffers, the device part
        of the step is captured once (after `graph_warmup` eager steps) and replayed afterwards."""
        if not self.use_graph:
            x_real = self.converter(x_real, self.device)
            t_real = None if t_real is None else self.converter(t_real, self.device)
            self._report_copied()
            return self.step_on_device(x_real, t_real)
        as_t = lambda a: a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a))
        x_real = as_t(x_real)
        t_real = None if t_real is None else as_t(t_real)
        if self._static is None:
            dev = torch.device("cuda", torch.cuda.current_device())
            self._static = (torch.empty_strided(x_real.shape, x_real.stride(), dtype=x_real.dtype, device=dev),
                            None if t_real is None else torch.empty(t_real.shape, dtype=torch.int32, device=dev))
        sx, st = self._static
        sx.copy_(x_real, non_blocking=True)
        if st is not None:
            st.copy_(t_real, non_blocking=True)
        self._report_copied()
        if self._graph is not None:
            self._graph.replay()
            self._report_replayed()
            return
        if self._eager_steps < self.graph_warmup:
            self._eager_steps += 1
            return self.step_on_device(sx, st)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.step_on_device(sx, st)
        self._graph = g
        g.replay()
        self._report_replayed()


def _lab(t):
    if t is None:
        return None
    d = t.data if isinstance(t, Variable) else t
    return d if d.dtype == torch.int32 else d.int()
