"""BaseModel — host-side mirror of /root/reference/models/basemodel.py.

Same constructor kwargs, modes ('TRAINING' / 'INFERENCE'), methods
(`train_step`, `infer`, `test`, `snapshot`) and attributes (`y_hat`,
`y_hat_sig`, `output`, `inference_ops`, `seg_loss_op`, `global_step`,
`batch_size`, `n_kernels`, `model_name`) as the reference, but `sess.run` is
replaced by an explicit kernel schedule over the C ABI (include/segb200.h),
captured into a CUDA graph.  Evident intent is implemented where the
reference is broken at HEAD (SURVEY.md §7 "Reference defects"):

  * `train_step()` (raises at reference `basemodel.py:477-478`; intended body
    in the comments at `:480-489`): one Adam step on
    mean softmax-cross-entropy (`:59-70`, `:194`, `:360`), `global_step += 1`.
  * `infer(imgs)` (`:527-531`): `[sigmoid(y_hat), float32(argmax(...)[..., None])]`.
  * INFERENCE mode accepts `dataset=None` (reference dereferences it at `:39`).

Dataset duck-type consumed (reference `:39,95,159-169`): `.batch_size`,
`.use_feed`, `.has_masks`, `.set_tf_sess(sess)`; the TF queue ops
`.image_op/.mask_op` become `.next_batch()` -> (images fp32 [B,H,W,C] in
[0,1], masks uint8 [B,H,W,1]) as numpy arrays or torch tensors.
"""
import os

import numpy as np
import torch

from .. import engine as E
from .. import native as N


class BaseModel(object):
    def __init__(self,
                 sess,
                 mode='TRAINING',
                 log_dir='./logs',
                 dataset=None,
                 test_dataset=None,
                 bayesian=False,
                 save_dir='./snapshot',
                 n_classes=None,
                 input_dims=None,
                 input_channel=3,
                 autoencoder=False,
                 load_snapshot=True,
                 learning_rate=1e-3,
                 load_snapshot_from=None,
                 adversarial_training=False):
        self.mode = mode
        self.log_dir = log_dir
        self.dataset = dataset
        self.test_dataset = test_dataset
        self.save_dir = save_dir
        self.bayesian = bayesian
        self.n_classes = n_classes
        if isinstance(input_dims, int):          # reference default is an int (unet.py:32)
            input_dims = [input_dims, input_dims]
        self.input_dims = list(input_dims)
        self.autoencoder = autoencoder
        self.learning_rate = learning_rate
        self.input_channel = input_channel
        self.adversarial_training = adversarial_training
        if adversarial_training:
            raise Exception('adversarial_training is outside the B200 hot path (SURVEY.md §8 N3)')
        if mode not in ('TRAINING', 'INFERENCE'):
            raise Exception('mode must be TRAINING or INFERENCE')
        self.batch_size = self.dataset.batch_size if self.dataset is not None else None
        if mode == 'TRAINING' and self.dataset is None:
            raise Exception('TRAINING mode needs a dataset')

        self.IN_OUT_EQUAL = False
        self.IN_OUT_CROP = False
        self.IN_OUT_RATIO = False

        self.load_snapshot = load_snapshot if load_snapshot else False
        if self.mode == 'INFERENCE':
            self.load_snapshot = True
        self.load_snapshot_from = load_snapshot_from if load_snapshot_from else False
        self.summary_iter = 25

        # ---- device / library
        self.sess = sess                          # opaque, ignored
        N.load()
        N.check(N.load().seg_device_check(), 'seg_device_check')
        self.device = torch.device('cuda', torch.cuda.current_device())
        self.impl = N.IMPL_SIMT if os.environ.get('SEGB200_IMPL', 'umma') == 'simt' else N.IMPL_UMMA
        self.global_step = 0
        self.store = E.ParamStore(self.device)
        self._exec = {}                           # (B, training) -> executor
        self._loss_sum = None
        self._last_pixels = 1
        self.mc_seed = 0
        # global_step as a device scalar: kernels whose behaviour follows the step (dropout
        # streams) read it at run time, so captured CUDA graphs stay valid across replays
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._last_train_exec = None              # executor of the most recent train_step
        self._log_file = None
        self.world_size = 1
        self._allreduce = None                    # set by parallel.DataParallel: f(group index)
        self.opt_splits = ()                      # layers that start a new optimizer group
        if self.dataset is not None and hasattr(self.dataset, 'set_tf_sess'):
            self.dataset.set_tf_sess(self.sess)

    # ------------------------------------------------------------ child API
    def _build_layers(self, gen):
        raise NotImplementedError

    def _make_exec(self, batch, training):
        raise NotImplementedError

    def _finish_init(self, seed=0):
        """Called by the child constructor after it set its hyper-parameters:
        creates parameters (xavier-uniform weights, zero biases), loads a
        snapshot if asked to."""
        gen = np.random.default_rng(seed)
        self._build_layers(gen)
        self.store.finalize()
        for layer in self.layers.values():
            if hasattr(layer, 'init_values'):
                layer.init_values()
        self.store.refresh_shadow()
        self._init_opt_groups()
        self._init_saver(self.model_name)

    def _init_opt_groups(self):
        """Optimizer groups = contiguous slices of the flat parameter buffer (in TF variable
        order) that the backward pass completes one after the other, back to front.  Each
        group gets its own (data-parallel: all-reduce +) Adam launch as soon as its
        gradients are complete, beside the rest of the backward pass."""
        from ..parallel import bucket_boundaries
        st = self.store
        splits = [s for s in self.opt_splits if s + '/weights' in st.params]
        bounds = bucket_boundaries(st, splits)
        firsts = [None] + [s for s in splits if st.params[s + '/weights'].offset > 0]
        self.opt_groups = []
        for a, b, first in zip(bounds[:-1], bounds[1:], firsts):
            self.opt_groups.append({'first_layer': first, 'slice': (a, b),
                                    'chunks': st.chunk_range(a, b)})
        # group 0 starts with the first layer of the model, whatever its name
        self.group_of_first = {g['first_layer']: i for i, g in enumerate(self.opt_groups)
                               if g['first_layer'] is not None}

    # ----------------------------------------------------------- snapshots
    def _init_saver(self, name='model'):
        self.save_path = None
        if self.save_dir is not None:
            if not os.path.exists(self.save_dir):
                os.makedirs(self.save_dir)
            self.save_path = os.path.join(self.save_dir, '{}.ckpt'.format(name))
        if not self.load_snapshot:
            return
        try:
            path = self.load_snapshot_from if self.load_snapshot_from else self._latest_checkpoint()
            if path is None:
                raise IOError('no checkpoint')
            sd = dict(np.load(path, allow_pickle=False))
            # validate everything before touching the model (tf.train.Saver.restore checks
            # names and shapes before it assigns): a mismatched file leaves the freshly
            # initialised model, step counters included, exactly as it was
            st = self.store
            for name, p in st.params.items():
                if name not in sd:
                    raise KeyError('variable %s missing from %s' % (name, path))
                if int(np.prod(sd[name].shape)) != p.numel or \
                        tuple(sd[name].shape) not in (tuple(p.shape), (p.numel,)):
                    raise ValueError('variable %s: snapshot shape %s != %s' %
                                     (name, tuple(sd[name].shape), tuple(p.shape)))
            for name, t in st.state.items():
                if name in sd and tuple(sd[name].shape) != tuple(t.shape):
                    raise ValueError('state %s: snapshot shape %s != %s' %
                                     (name, tuple(sd[name].shape), tuple(t.shape)))
            for k in ('__flat__/segAdam', '__flat__/segAdam_1'):
                if k in sd and sd[k].shape != (st.numel,):
                    raise ValueError('%s: %s != (%d,)' % (k, sd[k].shape, st.numel))
            gs = int(sd.pop('global_step', 0))
            adam_step = int(sd.pop('adam_step', gs))
        except (IOError, OSError, KeyError, ValueError) as e:
            print('Failed to load snapshot; proceed with training')
            self._snapshot_error = e
            return
        self.global_step, st.step = gs, adam_step
        st.load_state_dict(sd)
        for k, buf in (('segAdam', st.m), ('segAdam_1', st.v)):
            if ('__flat__/' + k) in sd:
                buf.copy_(torch.from_numpy(sd['__flat__/' + k]).to(self.device))
        self._sync_step_dev()
        print('Success! Resuming from global step {}'.format(self.global_step))

    def _latest_checkpoint(self):
        if self.save_dir is None or not os.path.isdir(self.save_dir):
            return None
        best, best_gs = None, -1
        for f in os.listdir(self.save_dir):
            if f.startswith(self.model_name + '.ckpt-') and f.endswith('.npz'):
                gs = int(f[len(self.model_name) + 6:-4])
                if gs > best_gs:
                    best, best_gs = os.path.join(self.save_dir, f), gs
        return best

    def snapshot(self):
        """`saver.save(sess, save_path, global_step=gs)` (reference :494-501);
        variables keep their TF names and layouts."""
        if self.mode == 'INFERENCE':
            print('snapshot() with INFERENCE mode invalid')
            return
        if self.save_dir is None or self.save_path is None:
            raise Exception('snapshot() needs save_dir (the model was built with save_dir=None)')
        sd = self.store.state_dict()
        sd['global_step'] = np.int64(self.global_step)
        sd['adam_step'] = np.int64(self.store.step)
        sd['__flat__/segAdam'] = self.store.m.cpu().numpy()
        sd['__flat__/segAdam_1'] = self.store.v.cpu().numpy()
        path = '{}-{}.npz'.format(self.save_path, self.global_step)
        np.savez(path, **sd)
        for f in os.listdir(self.save_dir):       # max_to_keep=1
            full = os.path.join(self.save_dir, f)
            if f.startswith(self.model_name + '.ckpt-') and full != path:
                os.remove(full)
        print('Global step {}, snapshotting to {}'.format(self.global_step, path))
        return path

    # -------------------------------------------------------------- helpers
    def _get_exec(self, batch, training):
        key = (batch, training)
        if key not in self._exec:
            self._exec[key] = self._make_exec(batch, training)
        return self._exec[key]

    def _to_device(self, arr, dtype):
        if isinstance(arr, np.ndarray):
            arr = torch.from_numpy(np.ascontiguousarray(arr))
        return arr.to(self.device, dtype, non_blocking=True).contiguous()

    def _sync_step_dev(self):
        self.step_dev.fill_(self.global_step)

    def _log_step(self):
        """Every `summary_iter` steps: one JSON line {"step", "loss"} appended to
        <log_dir>/train_log.jsonl (the reference writes a TF summary at that cadence,
        models/basemodel.py:74,486-489)."""
        if self.log_dir is None or self.global_step % self.summary_iter != 0:
            return
        if self._log_file is None:
            if not os.path.exists(self.log_dir):
                os.makedirs(self.log_dir)
            self._log_file = open(os.path.join(self.log_dir, 'train_log.jsonl'), 'a')
        import json
        self._log_file.write(json.dumps({'step': self.global_step, 'loss': self.seg_loss_op}) + '\n')
        self._log_file.flush()

    def load_weights(self, state_dict):
        """Inject parameters given under their TF variable names / layouts."""
        self.store.load_state_dict(state_dict)

    # ----------------------------------------------------------- train_step
    def train_step(self, batch=None):
        """One optimizer step (`sess.run(self.train_op_list)`, reference
        :484/:488): forward, mean softmax x-entropy, backward, Adam, global_step+1."""
        if self.mode == 'INFERENCE':
            raise Exception('train_step() with INFERENCE mode invalid')
        if batch is None:
            # the reference's call: the batch comes from the dataset.  The NEXT batch's
            # host->device copy is issued on a copy stream right behind this step's
            # launch, so it overlaps the step's kernels (the reference's queue runners
            # prefetch the same way, utils/datasets.py:94-196).
            ex = self._get_exec(self.dataset.batch_size, True)
            self._last_train_exec = ex
            ex.train_step_from(self.dataset)
            self.global_step += 1
            self._log_step()
            return
        imgs, masks = batch[0], batch[1]
        crop = batch[2] if len(batch) > 2 else None      # uint8 batches: per-image crop corners
        # host (pinned or pageable) or device tensors: staged straight into the
        # executor's static input buffers (H2D on the compute stream)
        x = torch.from_numpy(imgs) if isinstance(imgs, np.ndarray) else imgs
        y = torch.from_numpy(masks) if isinstance(masks, np.ndarray) else masks
        if isinstance(crop, np.ndarray):
            crop = torch.from_numpy(crop)
        ex = self._get_exec(x.shape[0], True)
        self._last_train_exec = ex
        ex.train_step(x, y, crop)
        self.global_step += 1
        self._log_step()

    @property
    def seg_loss_op(self):
        """Mean cross-entropy of the most recent train_step (host float)."""
        ex = self._last_train_exec
        if ex is None:
            return float('nan')
        return ex.loss_value(0)

    @property
    def seg_loss_lagged(self):
        """Mean cross-entropy of the step BEFORE the most recent train_step: a training
        loop that logs this keeps one step in flight on the GPU instead of draining the
        stream after every step (every step's loss is still copied to the host)."""
        ex = self._last_train_exec
        if ex is None:
            return float('nan')
        return ex.loss_value(1)

    # ---------------------------------------------------------------- infer
    def infer(self, imgs):
        """`sess.run(self.inference_ops, {input_x: imgs})` (reference :527-531).
        imgs: fp32 [B,H,W,C] in [0,1].  Returns [probs [B,H',W',n_classes],
        labelmap [B,H',W',1]] as float32 numpy arrays."""
        x = self._to_device(imgs, torch.float32)
        ex = self._get_exec(x.shape[0], False)
        self._sync_step_dev()
        probs, labels = ex.infer(x)
        torch.cuda.current_stream().synchronize()
        return [probs.cpu().numpy(), labels.cpu().numpy()]

    def test(self):
        """Loss on one batch of `test_dataset` through the weight-sharing tower
        (reference :506-518, :397-436) without touching the parameters."""
        if self.mode == 'INFERENCE':
            print('test() with INFERENCE mode invalid')
            return None
        ds = self.test_dataset if self.test_dataset is not None else self.dataset
        imgs, masks = ds.next_batch()
        x = self._to_device(imgs, torch.float32)
        y = self._to_device(masks, torch.uint8)
        ex = self._get_exec(x.shape[0], False)
        self._sync_step_dev()
        loss = ex.eval_loss(x, y)
        print('TEST LOSS', loss, self.global_step)
        return loss


class ExecBase(object):
    """Buffers + kernel schedule of one model for one (batch size, mode).
    Children allocate activations / gradients and implement forward() and
    backward(); this base owns input staging, loss, head, the CUDA-graph
    captured train step and inference."""

    @staticmethod
    def first_layer_x4(model, layer):
        """True if `layer` (the model's first convolution) can run on the first-layer kernel:
        a plain 3x3 stride-1 convolution of an RGB input onto 32 or 64 padded channels
        (csrc/fconv.cuh).  SEGB200_FIRST_LAYER=0 keeps the 16-channel padded input."""
        return (type(layer) is E.ConvLayer and layer.kind == 'conv' and layer.k == 3 and
                layer.stride == 1 and model.input_channel == 3 and layer.cout_pad in (32, 64) and
                os.environ.get('SEGB200_FIRST_LAYER', '1') != '0')

    def _init_io(self, model, B, H, W, oh, ow, n_out, dlogits_pad, training):
        dev = model.device
        self.m, self.B, self.training = model, B, training
        self.H, self.W, self.oh, self.ow = H, W, oh, ow
        # x_f32 / mask_in are the host->device landing buffers; the packed bf16 input
        # (act['x']) and `mask` are what the kernels of a step read, so the next batch
        # may land while a step is running
        self.x_f32 = torch.zeros(B, H, W, model.input_channel, dtype=torch.float32, device=dev)
        self.mask_in = torch.zeros(B, H, W, 1, dtype=torch.uint8, device=dev)
        # x4: the executor's first layer takes the (R,G,B,1) input of seg_stage_input (set by
        # the child before _init_io).  One staging launch then does everything a step needs
        # in front of its graph: pack, mask, loss publication + zeroing, lr_t, global step.
        self.x4 = bool(getattr(self, 'x4', False))
        self.x_u8 = None               # uint8 landing buffer (raw-image batches), on demand
        self._src = (self.x_f32, self.mask_in, None)   # what the next staging launch reads
        self._loss_prezeroed = False
        self._loss_ring = torch.zeros(8, dtype=torch.float32).pin_memory()
        self._stage_ev = [torch.cuda.Event(), torch.cuda.Event()]
        self.mask = torch.zeros(B, H, W, 1, dtype=torch.uint8, device=dev)
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.ev_staged = torch.cuda.Event()
        self.ev_consumed = torch.cuda.Event()
        self._pf = None                # host batch whose copy into the landing buffers is in flight
        self._pf_host = None           # host batch fetched from the dataset but not staged
        self._prepacked = False
        self.side = E.SideStream(dev, lanes=int(os.environ.get('SEGB200_WGRAD_LANES', '1')))   # weight gradients
        # Adam per optimizer group.  SEGB200_OPT_PRIO=1 makes it a high-priority stream (its
        # blocks are placed first when SM slots free up); measured: one GPU 0.963 vs 0.942 ms
        # (the prioritised Adam takes SMs from the critical-path kernels), 8 GPUs no gain once
        # the all-reduces have their own stream - off.
        self.opt = E.SideStream(dev, priority=-1 if os.environ.get('SEGB200_OPT_PRIO', '0') != '0'
                                else 0)
        # data parallel: the all-reduce of a group is issued from its own stream, so that the
        # next group's reduction does not queue behind this group's Adam (8-GPU timeline: the
        # conv3-4 bucket was ready at ~930 us and started at 1013 us, behind the bottleneck
        # group's Adam; 1.140 -> 1.080 ms/step, profiles/r02_dp.md)
        self.comm = E.SideStream(dev)
        self.skipside = E.SideStream(dev)         # skip-connection halves of the concat input gradients
        self.use_side = os.environ.get('SEGB200_WGRAD_STREAM', '1') != '0'
        self._opt_active = False
        self._pending = set()
        self.logits = torch.zeros(B, oh, ow, n_out, dtype=torch.float32, device=dev)
        self.probs = torch.zeros(B, oh, ow, n_out, dtype=torch.float32, device=dev)
        self.labelmap = torch.zeros(B, oh, ow, 1, dtype=torch.float32, device=dev)
        self.loss_sum = torch.zeros(1, dtype=torch.float32, device=dev)
        self._loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()
        self._loss_ev = [torch.cuda.Event(), torch.cuda.Event()]
        self.loss_pixels = B * oh * ow
        # mask centre crop when the logits are smaller than the input (unet.py:71-72)
        self.my0, self.mx0 = (H - oh) // 2, (W - ow) // 2
        self.dlogits = (torch.zeros(B, oh, ow, dlogits_pad, dtype=torch.bfloat16, device=dev)
                        if training else None)
        self.graph = None
        self.calls = 0
        self.use_graph = os.environ.get('SEGB200_NO_GRAPH', '0') != '1'

    # ------------------------------------------------------------- staging
    def stage(self, x, mask):
        self.x_f32.copy_(x, non_blocking=True)
        if mask is not None:
            self.mask.copy_(mask, non_blocking=True)

    def mask_view(self):
        return self.mask[:, self.my0:self.my0 + self.oh, self.mx0:self.mx0 + self.ow, :]

    def head(self):
        E.sigmoid_argmax(self.logits, self.probs, self.labelmap)
        self.m.y_hat_sig, self.m.output = self.probs, self.labelmap
        return self.probs, self.labelmap

    def zero_loss(self):
        """loss_sum = 0, unless the step's staging launch already did it."""
        if not self._loss_prezeroed:
            E.fill_zero(self.loss_sum)

    def forward_for_step(self):
        """The forward pass of a train step (executors whose step has a shorter route to the
        loss than forward() + loss() override this)."""
        self.forward()

    def loss(self, with_grad):
        self.zero_loss()
        E.softmax_xent(self.logits, self.mask_view(), self.loss_sum,
                       self.dlogits if with_grad else None)

    # ---------------------------------------------------------------- steps
    def _step_body(self):
        m = self.m
        self._opt_active = True
        self._sm_reserved = False
        self._pending = set(range(len(m.opt_groups)))
        try:
            self.forward_for_step()
            self.loss(True)
            self.backward()                       # may call group_ready() as groups complete
            for i in sorted(self._pending, reverse=True):
                self.group_ready(i)
            self.side.join()
            self.skipside.join()
            self.comm.join()
            self.opt.join()
        finally:
            self._opt_active = False
            if self._sm_reserved:
                N.set_option(N.OPT_SM_LIMIT, 0)
                self._sm_reserved = False

    def group_ready(self, i):
        """Every gradient of optimizer group i has been enqueued (weight gradients on the
        side stream; the input-gradient kernels that read the group's weights on the current
        stream).  On the optimizer stream, behind both: data-parallel all-reduce of the
        group's gradient slice, then Adam on it (which also rewrites the bf16 shadows and
        zeroes the gradients).  No-op outside a train step."""
        m = self.m
        if not self._opt_active or i not in self._pending:
            return
        self._pending.discard(i)
        split = m._allreduce is not None and os.environ.get('SEGB200_COMM_STREAM', '1') != '0'
        reserve = int(os.environ.get('SEGB200_DP_SM_RESERVE', '0'))
        if m._allreduce is not None and reserve > 0 and not self._sm_reserved:
            # from the first bucket on, the tile kernels' persistent grids leave `reserve` SMs
            # to the all-reduce kernels (grid sizes are fixed at graph capture)
            sms = torch.cuda.get_device_properties(m.device).multi_processor_count
            N.set_option(N.OPT_SM_LIMIT, sms - reserve)
            self._sm_reserved = True
        if split:
            with self.comm.fork(also=(self.side, self.skipside)):
                m._allreduce(i)
        with self.opt.fork(also=(self.side, self.skipside, self.comm) if split
                           else (self.side, self.skipside)):
            if m._allreduce is not None and not split:
                m._allreduce(i)
            m.store.adam_launch(0.0, grad_scale=1.0 / m.world_size, from_device=True,
                                chunk_range=m.opt_groups[i]['chunks'])

    def layer_done(self, name):
        """Backward schedules call this after a layer's backward kernels were enqueued."""
        gi = self.m.group_of_first.get(name)
        if gi is not None:
            self.group_ready(gi)

    def pack(self):
        """fp32 [B,H,W,C] landing buffer -> bf16 input tensor of the first layer."""
        if not self._prepacked:
            self._pack_now()

    def _pack_now(self):
        """Default: the first-layer kernel's (R,G,B,1) layout when the executor uses it,
        else C zero-padded to 16 channels.  Executors whose first layer is a PatchConvLayer
        override this with its patch packing."""
        if self.x4:
            E.stage_input(self.x_f32, self.act['x'])
        else:
            E.pack_input(self.x_f32, self.act['x'])

    def _stage_async(self, x, mask, crop_yx=None):
        """Make a batch available to the next step.  Host tensors are copied into the landing
        buffers over the copy stream (pinned memory makes this asynchronous); device tensors
        go over the compute stream, or - x4 executors, contiguous fp32 / uint8 - are read in
        place by the step's staging launch.  uint8 images (raw decoded files, reference
        utils/datasets.py:19-45,176-190) may be larger than the model input: `crop_yx`
        (int32 [B,2]) then gives each image's crop corner, and the masks are raw 0/255."""
        cur = torch.cuda.current_stream()
        raw = x.dtype == torch.uint8
        if raw or crop_yx is not None:
            if not self.x4:
                raise Exception('uint8 / cropped batches need the first-layer (x4) input path')
        if self.x4 and x.is_cuda and x.is_contiguous() and mask.is_cuda and mask.is_contiguous() \
                and mask.dtype == torch.uint8 and (raw or x.dtype == torch.float32) \
                and (crop_yx is not None or tuple(x.shape[1:3]) == (self.H, self.W)):
            cy = None if crop_yx is None else crop_yx.to(device=x.device, dtype=torch.int32)
            cur.wait_event(self.ev_staged)
            self._src = (x, mask, cy)
            self.ev_staged.record(cur)
            return
        if raw:
            if self.x_u8 is None or tuple(self.x_u8.shape) != tuple(x.shape):
                self.x_u8 = torch.zeros(tuple(x.shape), dtype=torch.uint8, device=self.m.device)
                self.mask_u8 = torch.zeros(tuple(mask.shape), dtype=torch.uint8,
                                           device=self.m.device)
                self.crop_dev = torch.zeros(x.shape[0], 2, dtype=torch.int32, device=self.m.device)
            land_x, land_m = self.x_u8, self.mask_u8
        else:
            land_x, land_m = self.x_f32, self.mask_in
        if x.is_cuda:
            cur.wait_event(self.ev_staged)           # a prefetch still landing there
            land_x.copy_(x, non_blocking=True)
            land_m.copy_(mask, non_blocking=True)
            if raw and crop_yx is not None:
                self.crop_dev.copy_(crop_yx, non_blocking=True)
            self.ev_staged.record(cur)
        else:
            cs = self.copy_stream
            cs.wait_event(self.ev_consumed)              # previous batch packed
            with torch.cuda.stream(cs):
                land_x.copy_(x, non_blocking=True)
                land_m.copy_(mask, non_blocking=True)
                if raw and crop_yx is not None:
                    self.crop_dev.copy_(crop_yx, non_blocking=True)
                self.ev_staged.record(cs)
        self._src = (land_x, land_m, self.crop_dev if (raw and crop_yx is not None) else None)

    def _prefetch(self, dataset):
        batch = self._pf_host if self._pf_host is not None else dataset.next_batch()
        self._pf_host = None
        imgs, masks = batch[0], batch[1]
        crop = batch[2] if len(batch) > 2 else None
        x = torch.from_numpy(imgs) if isinstance(imgs, np.ndarray) else imgs
        y = torch.from_numpy(masks) if isinstance(masks, np.ndarray) else masks
        if isinstance(crop, np.ndarray):
            crop = torch.from_numpy(crop)
        self._stage_async(x, y, crop)
        self._pf = batch

    def train_step_from(self, dataset):
        """One step on the dataset's next batch; the following batch's H2D copy is
        issued behind this step's kernels."""
        if self._pf is None:
            self._prefetch(dataset)
        self._pf = None
        self._launch_step()
        self._prefetch(dataset)

    def train_step(self, x, mask, crop_yx=None):
        if self._pf is not None:                     # keep the prefetched batch for later
            self._pf_host, self._pf = self._pf, None
        self._stage_async(x, mask, crop_yx)
        self._launch_step()

    def _launch_step(self):
        m = self.m
        cur = torch.cuda.current_stream()
        cur.wait_event(self.ev_staged)
        lr_t = m.store.next_lr_t(m.learning_rate)
        if self.x4:
            # ONE launch in front of the step's graph: input pack (+ /255, crop), mask, the
            # previous step's loss to the pinned ring, loss_sum = 0, lr_t and global step
            x, mask, crop = self._src
            ctl = N.SegStageCtl(self.loss_sum.data_ptr(), self._loss_ring.data_ptr(),
                                self.calls - 1, float(lr_t), m.store.lr_t_dev.data_ptr(),
                                int(m.global_step), m.step_dev.data_ptr())
            E.stage_input(x, self.act['x'], mask_src=mask, mask_dst=self.mask, crop_yx=crop,
                          ctl=ctl)
            self._stage_ev[self.calls & 1].record(cur)
            self._src = (self.x_f32, self.mask_in, None)
            self._loss_prezeroed = True
        else:
            self._pack_now()
            self.mask.copy_(self.mask_in, non_blocking=True)
            m.store.lr_t_dev.fill_(lr_t)
            m._sync_step_dev()
        self.ev_consumed.record(cur)
        self._prepacked = True
        try:
            self._run_step()
        finally:
            self._prepacked = False
            self._loss_prezeroed = False

    def _run_step(self):
        m = self.m
        if self.use_graph and self.graph is None and self.calls >= 1:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            # Optional (SEGB200_MAIN_PRIO=1): capture the critical path on a stream of higher
            # priority than the side / optimizer streams.  Measured worse (1.456 vs 1.247
            # ms/step): the starved weight-gradient kernels pile up behind the backward pass.
            prio = os.environ.get('SEGB200_MAIN_PRIO', '0') == '1'
            cap = torch.cuda.Stream(device=m.device, priority=-1) if prio else None
            with torch.cuda.graph(g, stream=cap):
                self._step_body()
            self.graph = g
        if self.graph is not None:
            self.graph.replay()
        else:
            self._step_body()
        if not self.x4:
            # the step's loss goes to pinned host memory behind the step (4 bytes,
            # asynchronous): the host can read step i's loss while step i+1 runs
            # (BaseModel.seg_loss_lagged).  x4 executors: the NEXT step's staging launch
            # publishes it, no node behind the graph at all.
            slot = self.calls & 1
            self._loss_host[slot].copy_(self.loss_sum[0], non_blocking=True)
            self._loss_ev[slot].record(torch.cuda.current_stream())
        self.calls += 1

    def loss_value(self, lag=0):
        """Mean loss of the most recent step (lag=0) or of the one before it (lag=1).
        x4 executors: step n's loss is published to the pinned ring by the staging launch of
        step n+1 (the latest step's is read from the device after a stream sync); otherwise
        it is read from the pinned copy once that step's event has completed."""
        if self.calls == 0:
            return float('nan')
        n = self.calls - 1 - (lag if self.calls > lag else 0)
        if not self.x4:
            self._loss_ev[n & 1].synchronize()
            return float(self._loss_host[n & 1]) / self.loss_pixels
        if n == self.calls - 1:
            return float(self.loss_sum.item()) / self.loss_pixels
        self._stage_ev[(n + 1) & 1].synchronize()
        slot = 2 * (n & 3)
        if int(self._loss_ring.view(torch.int32)[slot + 1]) != n:
            raise Exception('loss ring: slot %d does not hold step %d' % (slot // 2, n))
        return float(self._loss_ring[slot]) / self.loss_pixels

    def infer(self, x):
        self.stage(x, None)
        self.forward()
        return self.head()

    def eval_loss(self, x, mask):
        self.stage(x, mask)
        self.forward()
        E.fill_zero(self.loss_sum)
        E.softmax_xent(self.logits, self.mask_view(), self.loss_sum, None)
        return float(self.loss_sum.item()) / self.loss_pixels
