"""The frame-sharded inference hot path as one callable: backbone features in, 2D + 3D joints out.

    mvfex refinement (H1 D1 Q1 M1 F1 A1-3 T1 R1 H2)  ->  decode of the refined heatmap (D1)
    -> pose3d lifting (P1-P4)  ->  packed [B, V*15*2 + 16*3] fp32 row per frame (the all-gather payload)

Used by bench.py, __graft_entry__.smoke() and the multi-GPU harness (egorear_b200/dist.py).
"""
import torch

from . import ops, synth
from .configs import heatmap_mvfex_cfg, pose3d_cfg
from .modules import EgoPoseFormerHeatmapMVFEX, EgoPoseFormerPose3D


class HotPathPipeline:
    def __init__(self, num_views=4, camera_model="ego4view_syn", precision="fp16", device="cuda", synthetic_weights=True,
                 with_backbone=False, materialize_features=True, backbone_impl="egr"):
        """backbone_impl: "egr" = the backbone engine (tcgen05 conv stages) behind backbone_staged(); "torch" = the PyTorch
        modules under autocast (round 1)"""
        self.V, self.camera_model, self.precision = num_views, camera_model, precision
        self.backbone_impl = backbone_impl
        self.heatmap = EgoPoseFormerHeatmapMVFEX(**heatmap_mvfex_cfg(num_views, camera_model), precision=precision,
                                                 build_backbone=with_backbone)
        self.pose3d = EgoPoseFormerPose3D(**pose3d_cfg(num_views, camera_model), precision=precision)
        if synthetic_weights:
            synth.fill_state_dict(self.heatmap)
            synth.fill_state_dict(self.pose3d)
        self.heatmap = self.heatmap.to(device).eval()
        self.pose3d = self.pose3d.to(device).eval()
        # chained forward: pose3d reuses the channels-last copies, incl. a high-precision copy of the refined features in
        # the operand type of its proposal branch (fp16 by default in bf16 mode, else fp32 / TF32)
        pd = self.pose3d.engine().proposal_dtype()
        use_init = bool(getattr(self.pose3d, "use_pred_heatmap_init", True))
        hp = ("f16_only" if use_init else "f16") if pd == "f16" else "tf32" if pd == "tf32" else None
        self.heatmap.engine().export_staged(True, hp=hp)
        # materialize_features=False: like EgoPoseFormerMVFEX.forward (egoposeformer_mvf_ex.py:50-58), which returns poses
        # and heatmaps only, the refined features are not written out in NCHW fp32 (list_ff[1] is None)
        self.materialize_features = materialize_features
        self._lanes, self._next_lane, self._pending = None, 0, []

    def freeze(self):
        """weights will not change any more: skip the per-call parameter-version check"""
        for m in (self.heatmap, self.pose3d):
            m.engine()._sync_params()
            m.engine().frozen = True
        return self

    @torch.no_grad()
    def forward(self, feat, bfb, coord_trans_mat=None, heatmap_for_anchor=None, feat_staged=None, lane=0):
        """feat_staged: the features as backbone_staged() leaves them (view-major channels-last, 16-bit activation type);
        `feat` is then ignored and the hot path starts without its staging pass.  lane: which cached workspace pair the
        engines use (forwards that may be in flight at the same time on different streams need different lanes)."""
        list_hm, list_ff = self.heatmap.forward_from_feats(None if feat_staged is not None else feat, bfb, heatmap_for_anchor,
                                                           want_feat_refined=self.materialize_features,
                                                           feat_staged=feat_staged, lane=lane)
        B, V, J, H, W = list_hm[-1].shape
        pts2d, maxvals, valid = ops.get_max_preds(list_hm[-1].view(B * V, J, H, W), threshold=0.5, normalize=False)
        preds3d = self.pose3d(list_ff[0], list_ff[-1], list_hm[-1], coord_trans_mat, staged=self.heatmap, lane=lane)
        packed = ops.pack_joints(pts2d.view(B, V * J * 2), preds3d[-1])
        return dict(packed=packed, joints2d=pts2d.view(B, V, J, 2), pose3d=preds3d[-1], list_hm=list_hm, list_ff=list_ff,
                    list_pose3d=preds3d)

    __call__ = forward

    @torch.no_grad()
    def backbone(self, img, chunk=128, autocast=True):
        """The PyTorch ResNet18+FPN backbones (out of the hot path's scope): img [B,V,3,256,256] ->
        (feat [B,V,128,64,64], bfb [B,V,512,8,8]) fp32.  bf16 autocast + frame chunks keep it off the critical memory path."""
        if not getattr(self, "_bb_cl", False):           # cuDNN's tensor-core kernels want channels-last weights
            for n in ("heatmap_estimator_stereo_front", "heatmap_estimator_stereo_back"):
                if hasattr(self.heatmap, n):
                    getattr(self.heatmap, n).to(memory_format=torch.channels_last)
            self._bb_cl = True
        B = img.shape[0]
        feat = torch.empty((B, self.V, 128, 64, 64), dtype=torch.float32, device=img.device)
        bfb = torch.empty((B, self.V, 512, 8, 8), dtype=torch.float32, device=img.device)
        for i in range(0, B, chunk):
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                f, bb = self.heatmap.forward_heatmap_feat_estimation(img[i:i + chunk])
            feat[i:i + chunk].copy_(f)                   # cast + back to the NCHW layout of the module interface
            bfb[i:i + chunk].copy_(bb[-1])
        return feat, bfb

    @torch.no_grad()
    def backbone_staged(self, img, chunk=128):
        """backbone() for a chained forward (second half): the FPN output of a channels-last bf16 backbone
        already IS the layout the hot path reads, so it is handed over as such instead of being cast back to NCHW fp32
        and re-staged.  The two stereo backbones run on view-major batches (frames are independent: eval-mode BatchNorm).
        -> (feat_staged bf16 [V,B,64,64,128] view-major channels-last, bfb fp32 [B,V,512,8,8])"""
        from .engine import ACT_DTYPE
        assert self.V == 4 and self.precision in ACT_DTYPE
        if self.backbone_impl == "egr":
            B = img.shape[0]
            if B <= chunk * 4:
                return self.heatmap.forward_backbone_staged(img)
            xh = torch.empty((self.V, B, 64, 64, 128), dtype=ACT_DTYPE[self.precision], device=img.device)
            bfb = torch.empty((B, self.V, 512, 8, 8), dtype=torch.float32, device=img.device)
            for i in range(0, B, chunk * 4):
                f, b = self.heatmap.forward_backbone_staged(img[i:i + chunk * 4])
                xh[:, i:i + chunk * 4].copy_(f)
                bfb[i:i + chunk * 4].copy_(b)
            return xh, bfb
        adt = ACT_DTYPE[self.precision]
        if not getattr(self, "_bb_cl", False):
            for n in ("heatmap_estimator_stereo_front", "heatmap_estimator_stereo_back"):
                getattr(self.heatmap, n).to(memory_format=torch.channels_last)
            self._bb_cl = True
        B = img.shape[0]
        xh = torch.empty((self.V, B, 64, 64, 128), dtype=adt, device=img.device)
        bfb = torch.empty((B, self.V, 512, 8, 8), dtype=torch.float32, device=img.device)
        for p, name in enumerate(("heatmap_estimator_stereo_front", "heatmap_estimator_stereo_back")):
            est = getattr(self.heatmap, name)
            for i in range(0, B, chunk):
                im = img[i:i + chunk, 2 * p:2 * p + 2].transpose(0, 1)             # [2, b, 3, H, W]: view-major batch
                with torch.autocast("cuda", dtype=adt):
                    f, bb = est.forward_backbone(im)                              # f [2, b, 128, 64, 64], NHWC memory
                xh[2 * p:2 * p + 2, i:i + chunk].copy_(f.permute(0, 1, 3, 4, 2))     # same memory order: a plain copy
                bfb[i:i + chunk, 2 * p:2 * p + 2].copy_(bb[-1].transpose(0, 1))
        return xh, bfb

    # ---- latency mode: the whole forward as one CUDA graph ----
    def capture(self, B, with_coord_trans_mat=False):
        """Capture the synchronous forward for batch size B into a CUDA graph (the ~95 launches of a step, with their
        programmatic-dependent-launch edges, become one graph launch: what a real-time B=1 loop needs).  The forward
        allocates nothing on the device side (workspaces are cached per lane, split-K scratch lives in the workspace), so
        the capture is legal; weights must be frozen.  Use replay(feat, bfb) afterwards."""
        dev = next(self.heatmap.parameters()).device
        self.freeze()
        lane = 1000 + B                                    # private workspaces: replays may interleave with eager calls
        feat = torch.zeros((B, self.V, 128, 64, 64), dtype=torch.float32, device=dev)
        bfb = torch.zeros((B, self.V, 512, 8, 8), dtype=torch.float32, device=dev)
        ctm = torch.eye(4, device=dev).repeat(B, self.V, 1, 1).contiguous() if with_coord_trans_mat else None
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):                      # warm-up outside the capture: workspaces, lazy inits
            for _ in range(2):
                self.forward(feat, bfb, ctm, lane=lane)
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = self.forward(feat, bfb, ctm, lane=lane)
        if not hasattr(self, "_graphs"):
            self._graphs = {}
        self._graphs[B] = (graph, feat, bfb, ctm, out)
        return self

    def replay(self, feat, bfb, coord_trans_mat=None):
        """run the captured graph of this batch size on new inputs; returns the graph's static output tensors"""
        graph, sf, sb, sc, out = self._graphs[feat.shape[0]]
        sf.copy_(feat, non_blocking=True)
        sb.copy_(bfb, non_blocking=True)
        if sc is not None and coord_trans_mat is not None:
            sc.copy_(coord_trans_mat, non_blocking=True)
        graph.replay()
        return out

    # ---- throughput mode: independent batches on alternating streams ----
    def forward_async(self, feat, bfb, coord_trans_mat=None, heatmap_for_anchor=None, world=1, lanes=2):
        """Same computation as forward(), enqueued on one of `lanes` internal streams (round robin), each with its own
        workspace.  Batches are independent, so the latency-bound token phases of one batch (chains of small kernels that
        leave most SMs idle) overlap with the dense stages of the next one.  The results belong to the lane's stream:
        call wait(out) before consuming them on the current stream, or join() to wait for everything in flight.
        With world > 1 the packed joints are all-gathered inside the lane (out["gathered"])."""
        from . import dist as egd
        dev = feat.device
        cur = torch.cuda.current_stream(dev)
        if self._lanes is None or len(self._lanes) != lanes:
            self._lanes = [torch.cuda.Stream(dev) for _ in range(lanes)]
        k = self._next_lane
        self._next_lane = (k + 1) % lanes
        st = self._lanes[k]
        st.wait_stream(cur)                                # inputs were produced on the caller's stream
        with torch.cuda.stream(st):                        # lane 0 is the synchronous forward()
            out = self.forward(feat, bfb, coord_trans_mat, heatmap_for_anchor, lane=k + 1)
            out["gathered"] = egd.gather_rows(out["packed"], world)
            ev = torch.cuda.Event()
            ev.record(st)
        out["event"] = ev
        return out

    def wait(self, out):
        torch.cuda.current_stream().wait_event(out["event"])
        return out

    def join(self):
        """the current stream waits for every lane"""
        if self._lanes:
            cur = torch.cuda.current_stream()
            for st in self._lanes:
                cur.wait_stream(st)

    def stage_host_features(self, feat):
        """[B,V,128,64,64] fp32 -> [V,B,64,64,128] in the 16-bit activation type: the view-major channels-last layout a
        channels-last half-precision backbone leaves (backbone_staged) and the engines read without a staging pass; what
        infer_host_batches should be fed when the producer sits across PCIe (half the bytes of NCHW fp32)."""
        from .engine import ACT_DTYPE
        return feat.to(ACT_DTYPE[self.precision]).permute(1, 0, 3, 4, 2).contiguous()

    def stage_host_bottom(self, bfb):
        """the stride-32 map [B,V,512,8,8] in the 16-bit activation type for the trip across PCIe (it only feeds an 8x8 average
        pool + fc_bfb: the rounding is averaged over 64 values); infer_host_batches widens it back to fp32 on the device"""
        from .engine import ACT_DTYPE
        return bfb.to(ACT_DTYPE[self.precision]).contiguous()

    @torch.no_grad()
    def infer_host_batches(self, batches, world=1):
        """End-to-end serving loop over HOST batches: yields the packed joints of every batch as a CPU tensor.

        `batches` iterates (feat, bfb[, coord_trans_mat]) pinned-host tensors; feat is either fp32 [B,V,128,64,64] or the
        staged 16-bit [V,B,64,64,128] of stage_host_features (4.2 instead of 8.4 MB per frame over PCIe, and no staging
        pass on the device).  With a backbone (`with_backbone=True`) a batch may instead be (frames[, coord_trans_mat]) with
        `frames` the DECODED uint8 images [B,V,H,W,3] a data loader holds after PNG/JPEG decode (0.8 MB per 4-view frame at
        256x256): resize + ToTensor + Normalize (ops.preprocess_images, PIL-exact), the backbone engine and the hot path all
        run on the device, as the dataset-side drop-in (egorear_b200/datasets.py) arranges under the reference's run.py.
        The host->device copy of batch i+1 runs on a copy stream while batch i is computed (two device
        input slots, events both ways), so a step costs max(copy, compute) instead of their sum; the device->host read of
        the packed joints is the only synchronisation per batch.
        """
        from . import dist as egd
        dev = next(self.heatmap.parameters()).device
        main = torch.cuda.current_stream(dev)
        if getattr(self, "_copy_stream", None) is None:      # persistent: the allocator pools are per stream
            self._copy_stream = torch.cuda.Stream(dev)
            self._slots = [None, None]
        copy, slots = self._copy_stream, self._slots
        copy.wait_stream(main)
        copied = [torch.cuda.Event(), torch.cuda.Event()]
        consumed = [torch.cuda.Event(), torch.cuda.Event()]

        def upload(i, hb):
            s = i & 1
            with torch.cuda.stream(copy):
                if i >= 2:
                    copy.wait_event(consumed[s])         # the slot's previous batch has been read by the forward
                if slots[s] is None or len(slots[s]) != len(hb) or any(a.shape != b.shape or a.dtype != b.dtype for a, b in zip(slots[s], hb)):
                    slots[s] = tuple(torch.empty(t.shape, dtype=t.dtype, device=dev) for t in hb)
                for d, t in zip(slots[s], hb):
                    d.copy_(t, non_blocking=True)
                copied[s].record(copy)

        it = iter(batches)
        nxt = next(it, None)
        i = 0
        if nxt is not None:
            upload(0, nxt)
        while nxt is not None:
            cur_i, nxt = i, next(it, None)
            if nxt is not None:
                upload(cur_i + 1, nxt)                   # overlaps with the forward below
            s = cur_i & 1
            main.wait_event(copied[s])
            f = slots[s][0]
            if f.dtype == torch.uint8:                           # decoded frames: GPU preprocessing + backbone engine
                from . import ops
                ctm = slots[s][1] if len(slots[s]) > 1 else None
                xh, b = self.backbone_staged(ops.preprocess_images(f))
                packed = self.forward(None, b, ctm, feat_staged=xh)["packed"]
                consumed[s].record(main)
                yield egd.gather_rows(packed, world).cpu()
                i += 1
                continue
            b = slots[s][1]
            if b.dtype in (torch.bfloat16, torch.float16):       # stage_host_bottom: widened on the device
                b = b.float()
            ctm = slots[s][2] if len(slots[s]) > 2 else None
            if f.dtype in (torch.bfloat16, torch.float16):       # staged: read in place (also by the chained pose3d sampling)
                packed = self.forward(None, b, ctm, feat_staged=f)["packed"]
            else:
                packed = self.forward(f, b, ctm)["packed"]
            consumed[s].record(main)
            yield egd.gather_rows(packed, world).cpu()   # D2H of the step's result (synchronises the main stream)
            i += 1
