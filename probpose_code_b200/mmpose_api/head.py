"""``ProbMapHead`` with the reference's constructor, attribute names and ``state_dict`` layout
(mmpose/models/heads/hybrid_heads/probmap_head.py), inference side only.  ``forward`` /
``predict`` run on the sm_100a engine + fused decode kernel; the ``nn.Module`` children exist to
hold the weights under the reference's names (``deconv_layers.{0,1,3,4}``, ``final_layer``,
``{probability,visibility,oks,error}_layers.{0,1,4,5,8,9,12}``)."""
from __future__ import annotations

from typing import List, Sequence, Tuple, Union

import numpy as np
import torch
from torch import nn

from .. import ops
from ._engine_cache import EngineCache
from .registry import KEYPOINT_CODECS, MODELS, register
from .structures import InstanceData, PixelData, instance_data


class BaseHead(nn.Module):
    """mmpose/models/heads/base_head.py: ``decode`` dispatches to the codec."""

    decoder = None

    def forward(self, feats):
        raise NotImplementedError

    def predict(self, feats, batch_data_samples, test_cfg={}):
        raise NotImplementedError

    def loss(self, feats, batch_data_samples, train_cfg={}):
        raise NotImplementedError

    def decode(self, batch_outputs) -> List[InstanceData]:
        """base_head.py:33-86.  The GPU codec always supports batch decoding; a foreign codec
        without it is driven per instance on host copies exactly like the reference."""
        args = batch_outputs if isinstance(batch_outputs, tuple) else (batch_outputs,)
        if self.decoder is None:
            raise RuntimeError(f"The decoder has not been set in {self.__class__.__name__}. "
                               "Please set the decoder configs in the init parameters to "
                               "enable head methods `head.predict()` and `head.decode()`")
        if self.decoder.support_batch_decoding:
            batch_keypoints, batch_scores = self.decoder.batch_decode(*args)
        else:
            batch_keypoints, batch_scores = [], []
            for outputs in zip(*[a.detach().cpu().numpy() for a in args]):
                k, s = self.decoder.decode(*outputs)
                batch_keypoints.append(k)
                batch_scores.append(s)
        return [InstanceData(keypoints=k, keypoint_scores=s) for k, s in zip(batch_keypoints, batch_scores)]


def _scalar_branch(cin: int, cout: int, last: nn.Module) -> nn.Sequential:
    mods = []
    for ks in ((4, 3), (2, 2), (2, 2)):  # probmap_head.py:264: Conv -> BN -> MaxPool -> ReLU
        mods += [nn.Conv2d(cin, cin, 3, 1, 1), nn.BatchNorm2d(cin), nn.MaxPool2d(ks, ks), nn.ReLU(inplace=True)]
    mods += [nn.Conv2d(cin, cout, 1, 1, 0), last]
    return nn.Sequential(*mods)


_XY_COLS: dict = {}


def _record_xy(rec: np.ndarray) -> np.ndarray:
    """(B, K, F) records -> contiguous (B, K, 2) locations: one gather along the rows instead of a strided copy with
    numpy's inner loop over two elements (host time after the device has finished)."""
    b, k, f = rec.shape
    cols = _XY_COLS.get((k, f))
    if cols is None:
        cols = _XY_COLS[(k, f)] = (np.arange(k)[:, None] * f + np.arange(2)).reshape(-1)
    return np.take(rec.reshape(b, k * f), cols, axis=1).reshape(b, k, 2)


@register(MODELS, ["ProbMapHead"])
class ProbMapHead(BaseHead):
    _version = 2

    def __init__(self, in_channels: Union[int, Sequence[int]], out_channels: int,
                 deconv_out_channels=(256, 256, 256), deconv_kernel_sizes=(4, 4, 4), conv_out_channels=None,
                 conv_kernel_sizes=None, final_layer_dict: dict = dict(kernel_size=1), keypoint_loss=None,
                 probability_loss=None, visibility_loss=None, oks_loss=None, error_loss=None, normalize: float = None,
                 detach_probability: bool = True, detach_visibility: bool = True, learn_heatmaps_from_zeros: bool = False,
                 freeze_heatmaps: bool = False, freeze_probability: bool = False, freeze_visibility: bool = False,
                 freeze_oks: bool = False, freeze_error: bool = False,
                 decoder=dict(type="UDPHeatmap", input_size=(192, 256), heatmap_size=(48, 64), sigma=2), init_cfg=None,
                 precision: str = None):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.temperature = 0.5  # probmap_head.py:135
        self.normalize = normalize
        self.freeze_oks, self.freeze_error = freeze_oks, freeze_error
        self.test_cfg = {}
        # the five loss configs are accepted and ignored: training is out of scope here
        self.decoder = KEYPOINT_CODECS.build(decoder) if decoder is not None else None

        if deconv_out_channels:
            if deconv_kernel_sizes is None or len(deconv_out_channels) != len(deconv_kernel_sizes):
                raise ValueError('"deconv_out_channels" and "deconv_kernel_sizes" should '
                                 "be integer sequences with the same length. Got "
                                 f"mismatched lengths {deconv_out_channels} and {deconv_kernel_sizes}")
        unsupported = []
        if not isinstance(in_channels, int): unsupported.append("multi-level in_channels")
        if not deconv_out_channels or len(deconv_out_channels) != 2 or len(set(deconv_out_channels)) != 1:
            unsupported.append(f"deconv_out_channels={deconv_out_channels} (two equal deconv layers)")
        elif tuple(deconv_kernel_sizes) != (4, 4): unsupported.append(f"deconv_kernel_sizes={deconv_kernel_sizes} ((4, 4))")
        if conv_out_channels: unsupported.append("conv_out_channels")
        if final_layer_dict is None or dict(final_layer_dict).get("kernel_size", 1) != 1: unsupported.append("final layer != 1x1 conv")
        if normalize is None: unsupported.append("normalize=None (the sparsemax-normalised head only)")
        if unsupported:
            raise NotImplementedError("probpose_code_b200 ProbMapHead covers the shipped ProbPose configuration only; "
                                      "unsupported: " + ", ".join(unsupported))
        layers, c = [], in_channels
        for co in deconv_out_channels:  # probmap_head.py:441-470: k4 s2 p1 output_padding 0, no bias
            layers += [nn.ConvTranspose2d(c, co, 4, 2, 1, 0, bias=False), nn.BatchNorm2d(co), nn.ReLU(inplace=True)]
            c = co
        self.deconv_layers = nn.Sequential(*layers)
        self.conv_layers = nn.Identity()
        self.final_layer = nn.Conv2d(c, out_channels, 1)
        self.normalize_layer = nn.Identity()  # sparsemax has no parameters; it runs inside the decode kernel
        self.probability_layers = _scalar_branch(in_channels, out_channels, nn.Sigmoid())
        self.visibility_layers = _scalar_branch(in_channels, out_channels, nn.Sigmoid())
        self.oks_layers = _scalar_branch(in_channels, out_channels, nn.Sigmoid())
        self.error_layers = _scalar_branch(in_channels, out_channels, nn.ReLU())
        self.init_weights()
        self.eval()
        for p in self.parameters():
            p.requires_grad_(False)
        self._cache = EngineCache(dict(embed_dim=in_channels, depth=0, heads=0, ffn_dim=0, num_keypoints=out_channels,
                                       deconv_channels=deconv_out_channels[0], temperature=self.temperature,
                                       normalize=normalize), precision)
        self._named = None

    @property
    def default_init_cfg(self):
        return [dict(type="Normal", layer=["Conv2d", "ConvTranspose2d"], std=0.001),
                dict(type="Constant", layer="BatchNorm2d", val=1)]

    def init_weights(self):  # probmap_head.py:590-598
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
                nn.init.normal_(m.weight, std=0.001)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)

    def engine_tensors(self, prefix: str = "head."):
        if self._named is None or self._named[0] != prefix:
            named = {prefix + k: v for k, v in self.named_parameters()}
            named.update({prefix + k: v for k, v in self.named_buffers() if not k.endswith("num_batches_tracked")})
            self._named = (prefix, named)
        return self._named[1]

    # ---- forward ----------------------------------------------------------------------
    def _head_raw(self, x: torch.Tensor):
        eng = self._cache.get(self.engine_tensors(), x.shape[0], x.device)
        return eng.head(x.float().contiguous())

    @torch.no_grad()
    def forward(self, feats: Tuple[torch.Tensor]):
        """probmap_head.py:600-625: (heatmaps (B,K,H,W) sparsemax-normalised and clamped,
        probabilities, visibilities, oks, errors (B,K,1,1))."""
        x = feats[-1]
        logits, scal = self._head_raw(x)
        _, heatmaps = ops.decode(logits, input_is_logits=True, temperature=self.temperature, normalize=self.normalize,
                                 return_heatmaps=True)
        b, k = scal.shape[0], scal.shape[2]
        return (heatmaps,) + tuple(scal[:, j].reshape(b, k, 1, 1) for j in range(4))

    def forward_heatmap(self, x: torch.Tensor) -> torch.Tensor:
        return self.forward((x,))[0]

    def _codec(self):
        codec = self.decoder
        if codec is None or not hasattr(codec, "keypoints_from_locs"):
            raise RuntimeError(f"The decoder has not been set in {self.__class__.__name__} (ProbMap codec required)")
        return codec

    def alloc_records(self, batch: int):
        """The reference's per-person ``InstanceData`` (probmap_head.py:776-798) for a batch whose records have not
        arrived yet: fresh batch arrays plus one container per person holding VIEWS of them.  Everything here is host
        work that needs no result, so the fused path does it while the device is still computing; ``fill_records``
        then writes the batch arrays in place."""
        k = self.out_channels
        kpts = np.empty((batch, 1, k, 2), dtype=np.float64)
        cols = np.empty((5, batch, 1, k), dtype=np.float32)  # conf, prob, vis, oks, err
        conf, prob, vis, oks, err = cols
        scores = conf if self.freeze_oks else oks  # probmap_head.py:796-798
        preds = [instance_data(keypoints=a, keypoint_scores=s_, keypoints_conf=c, keypoints_probs=p, keypoints_visible=v,
                               keypoints_oks=o, keypoints_error=e)
                 for a, s_, c, p, v, o, e in zip(kpts, scores, conf, prob, vis, oks, err)]
        return (kpts, cols), preds

    def fill_records(self, arrays, rec: np.ndarray, to_image=None) -> None:
        """Host records (B, K, 7) -> the batch arrays of ``alloc_records``.  ``to_image``: optional callable mapping the
        whole batch's input-space keypoints (B, K, 2) to image space in one vectorised step (the estimator's
        topdown.py:165-167 arithmetic)."""
        kpts, cols = arrays
        k = self._codec().keypoints_from_locs(_record_xy(rec))
        kpts[:, 0] = k if to_image is None else to_image(k)
        cols[:, :, 0] = rec[:, :, 2:].transpose(2, 0, 1)

    def pack_records(self, records: torch.Tensor, to_image=None) -> List[InstanceData]:
        """Device records (B, K, 7) -> the reference's per-person ``InstanceData`` with ONE device->host copy for the
        batch; the per-person arrays are views of the batch arrays."""
        self._codec()
        rec = records.detach().cpu().numpy()
        arrays, preds = self.alloc_records(rec.shape[0])
        self.fill_records(arrays, rec, to_image)
        return preds

    @staticmethod
    def fused_test_cfg(test_cfg: dict) -> bool:
        """The fused kernels merge the flipped pass as ``flip_mode="heatmap", shift_heatmap=False`` (the shipped
        test_cfg); other settings run the flip merge as tensor ops (utils.flip_heatmaps) before the decode kernel."""
        return not test_cfg.get("flip_test", False) or (test_cfg.get("flip_mode", "heatmap") == "heatmap"
                                                        and not test_cfg.get("shift_heatmap", False))

    def fused_decoder(self) -> bool:
        """The fused sparsemax / TTA / expected-value kernel implements the ProbMap codec's ``"gaussian"`` decode; with any
        other ``decoder`` (the constructor default is UDPHeatmap, as in the reference) ``predict`` goes through
        ``BaseHead.decode`` -> that codec, as probmap_head.py:776 does."""
        from .codec import ProbMap
        return isinstance(self.decoder, ProbMap) and self.decoder.heatmap_type == "gaussian"

    def _predict_generic(self, feats, batch_data_samples, test_cfg: dict):
        """probmap_head.py:746-804 step by step on device tensors (any flip_mode / shift_heatmap / decoder)."""
        from .utils import flip_heatmaps
        if test_cfg.get("flip_test", False):
            assert isinstance(feats, list) and len(feats) == 2
            flip_indices = list(batch_data_samples[0].metainfo["flip_indices"])
            htm, *scal = self.forward(feats[0])
            htm_f, *scal_f = self.forward(feats[1])
            htm_f = flip_heatmaps(htm_f, flip_mode=test_cfg.get("flip_mode", "heatmap"), flip_indices=flip_indices,
                                  shift_heatmap=test_cfg.get("shift_heatmap", False))
            heatmaps = (htm + htm_f) * 0.5
            scal = [(a + b[:, flip_indices]) * 0.5 for a, b in zip(scal, scal_f)]
        else:
            heatmaps, *scal = self.forward(feats)
        b, c, h, w = heatmaps.shape
        preds = self.decode(heatmaps)
        prob, vis, oks, err = [s.detach().cpu().numpy().reshape((b, 1, c)) for s in scal]
        err = err / np.sqrt(h**2 + w**2)
        for pi, p in enumerate(preds):
            p.set_field(p["keypoint_scores"], "keypoints_conf")
            p.set_field(prob[pi], "keypoints_probs")
            p.set_field(vis[pi], "keypoints_visible")
            p.set_field(oks[pi], "keypoints_oks")
            p.set_field(err[pi], "keypoints_error")
            if not self.freeze_oks:
                p.set_field(oks[pi], "keypoint_scores")
        if test_cfg.get("output_heatmaps", False):
            return preds, [PixelData(heatmaps=hm) for hm in heatmaps.detach()]
        return preds

    @torch.no_grad()
    def predict(self, feats, batch_data_samples, test_cfg: dict = {}):
        """probmap_head.py:715-804.  Sparsemax, flip merge, decode and scalar merge run in one
        kernel over both passes' raw outputs."""
        if not (self.fused_test_cfg(test_cfg) and self.fused_decoder()):
            return self._predict_generic(feats, batch_data_samples, test_cfg)
        want_hm = bool(test_cfg.get("output_heatmaps", False))
        if test_cfg.get("flip_test", False):
            assert isinstance(feats, list) and len(feats) == 2
            flip_indices = batch_data_samples[0].metainfo["flip_indices"]
            x, xf = feats[0][-1], feats[1][-1]
            b = x.shape[0]
            logits, scal = self._head_raw(torch.cat([x, xf], 0))
            out = ops.decode(logits[:b], logits[b:], flip_indices, scal[:b], scal[b:], input_is_logits=True,
                             temperature=self.temperature, normalize=self.normalize, return_heatmaps=want_hm)
        else:
            logits, scal = self._head_raw(feats[-1])
            out = ops.decode(logits, scalars=scal, input_is_logits=True, temperature=self.temperature,
                             normalize=self.normalize, return_heatmaps=want_hm)
        if want_hm:
            records, heatmaps = out
            return self.pack_records(records), [PixelData(heatmaps=hm) for hm in heatmaps.detach()]
        return self.pack_records(out)

    def loss(self, feats, batch_data_samples, train_cfg: dict = {}):
        raise NotImplementedError("training (ProbMapHead.loss) is out of scope of the B200 inference path")


@register(MODELS, ["HeatmapHead"])
class HeatmapHead(BaseHead):
    """``HeatmapHead`` of the ViTPose td-hm configs (mmpose/models/heads/heatmap_heads/heatmap_head.py:22-268) with the
    reference's constructor and ``state_dict`` layout (``deconv_layers.{0,1,3,4}``, ``final_layer``), inference side
    only: the deconv stack runs on the same tcgen05 implicit-GEMM kernels as ProbMapHead's heatmap branch, flip merge
    and the UDPHeatmap (DARK-UDP) decode in one kernel (``pp_decode_udp``)."""

    _version = 2

    def __init__(self, in_channels: Union[int, Sequence[int]], out_channels: int, deconv_out_channels=(256, 256, 256),
                 deconv_kernel_sizes=(4, 4, 4), conv_out_channels=None, conv_kernel_sizes=None,
                 final_layer: dict = dict(kernel_size=1), loss=None, decoder=None, init_cfg=None, precision: str = None):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.test_cfg = {}
        self.decoder = KEYPOINT_CODECS.build(decoder) if decoder is not None else None
        if deconv_out_channels:
            if deconv_kernel_sizes is None or len(deconv_out_channels) != len(deconv_kernel_sizes):
                raise ValueError('"deconv_out_channels" and "deconv_kernel_sizes" should '
                                 "be integer sequences with the same length. Got "
                                 f"mismatched lengths {deconv_out_channels} and {deconv_kernel_sizes}")  # heatmap_head.py:83-88
        unsupported = []
        if not isinstance(in_channels, int): unsupported.append("multi-level in_channels")
        if not deconv_out_channels or len(deconv_out_channels) != 2 or len(set(deconv_out_channels)) != 1:
            unsupported.append(f"deconv_out_channels={deconv_out_channels} (two equal deconv layers)")
        elif tuple(deconv_kernel_sizes) != (4, 4): unsupported.append(f"deconv_kernel_sizes={deconv_kernel_sizes} ((4, 4))")
        if conv_out_channels: unsupported.append("conv_out_channels")
        if final_layer is None or dict(final_layer).get("kernel_size", 1) != 1: unsupported.append("final layer != 1x1 conv")
        if unsupported:
            raise NotImplementedError("probpose_code_b200 HeatmapHead covers the shipped ViTPose td-hm configuration only; "
                                      "unsupported: " + ", ".join(unsupported))
        layers, c = [], in_channels
        for co in deconv_out_channels:  # heatmap_head.py:141-173: k4 s2 p1 output_padding 0, no bias
            layers += [nn.ConvTranspose2d(c, co, 4, 2, 1, 0, bias=False), nn.BatchNorm2d(co), nn.ReLU(inplace=True)]
            c = co
        self.deconv_layers = nn.Sequential(*layers)
        self.conv_layers = nn.Identity()
        self.final_layer = nn.Conv2d(c, out_channels, 1)
        for m in self.modules():  # heatmap_head.py:189-195
            if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
                nn.init.normal_(m.weight, std=0.001)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
        self.eval()
        for p in self.parameters():
            p.requires_grad_(False)
        blur = getattr(self.decoder, "blur_kernel_size", 11)
        self._cache = EngineCache(dict(embed_dim=in_channels, depth=0, heads=0, ffn_dim=0, num_keypoints=out_channels,
                                       deconv_channels=deconv_out_channels[0], head_kind="heatmap", blur_kernel_size=blur),
                                  precision)
        self._named = None

    engine_tensors = ProbMapHead.engine_tensors
    fused_test_cfg = staticmethod(ProbMapHead.fused_test_cfg)

    def fused_decoder(self) -> bool:
        from .codec import UDPHeatmap
        return isinstance(self.decoder, UDPHeatmap) and self.decoder.heatmap_type == "gaussian"

    @torch.no_grad()
    def forward(self, feats: Tuple[torch.Tensor]) -> torch.Tensor:
        """heatmap_head.py:197-214: featmap -> heatmaps (B, K, 64, 48)."""
        x = feats[-1]
        eng = self._cache.get(self.engine_tensors(), x.shape[0], x.device)
        return eng.head(x.float().contiguous())

    def _codec(self):
        codec = self.decoder
        if codec is None or not hasattr(codec, "keypoints_from_locs"):
            raise RuntimeError(f"The decoder has not been set in {self.__class__.__name__} (UDPHeatmap codec required)")
        return codec

    def alloc_records(self, batch: int):
        """Per-person ``InstanceData(keypoints, keypoint_scores)`` (base_head.py:79-84) as views of fresh batch arrays;
        see :meth:`ProbMapHead.alloc_records`."""
        k = self.out_channels
        kpts = np.empty((batch, 1, k, 2), dtype=np.float64)
        scores = np.empty((batch, 1, k), dtype=np.float32)
        return (kpts, scores), [instance_data(keypoints=a, keypoint_scores=s_) for a, s_ in zip(kpts, scores)]

    def fill_records(self, arrays, rec: np.ndarray, to_image=None) -> None:
        kpts, scores = arrays
        k = self._codec().keypoints_from_locs(_record_xy(rec))
        kpts[:, 0] = k if to_image is None else to_image(k)
        scores[:, 0] = rec[:, :, 2]

    def pack_records(self, records: torch.Tensor, to_image=None) -> List[InstanceData]:
        """Device records (B, K, 3) -> per-person ``InstanceData``; ``to_image`` as in :meth:`ProbMapHead.fill_records`."""
        self._codec()
        rec = records.detach().cpu().numpy()
        arrays, preds = self.alloc_records(rec.shape[0])
        self.fill_records(arrays, rec, to_image)
        return preds

    @torch.no_grad()
    def predict(self, feats, batch_data_samples, test_cfg: dict = {}):
        """heatmap_head.py:216-265: flip merge + decode in one kernel over both passes' heatmaps."""
        if not (self.fused_test_cfg(test_cfg) and self.fused_decoder()):  # any other flip_mode / shift / codec: tensor ops
            from .utils import flip_heatmaps
            if test_cfg.get("flip_test", False):
                assert isinstance(feats, list) and len(feats) == 2
                hm = (self.forward(feats[0]) + flip_heatmaps(
                    self.forward(feats[1]), flip_mode=test_cfg.get("flip_mode", "heatmap"),
                    flip_indices=batch_data_samples[0].metainfo["flip_indices"],
                    shift_heatmap=test_cfg.get("shift_heatmap", False))) * 0.5
            else:
                hm = self.forward(feats)
            preds = self.decode(hm)
            return (preds, [PixelData(heatmaps=h) for h in hm.detach()]) if test_cfg.get("output_heatmaps", False) else preds
        want_hm = bool(test_cfg.get("output_heatmaps", False))
        blur = getattr(self.decoder, "blur_kernel_size", 11)
        if test_cfg.get("flip_test", False):
            assert isinstance(feats, list) and len(feats) == 2
            flip_indices = batch_data_samples[0].metainfo["flip_indices"]
            x, xf = feats[0][-1], feats[1][-1]
            b = x.shape[0]
            hm = self.forward((torch.cat([x, xf], 0),))
            out = ops.decode_udp(hm[:b], hm[b:], flip_indices, blur_kernel_size=blur, return_heatmaps=want_hm)
        else:
            out = ops.decode_udp(self.forward(feats), blur_kernel_size=blur, return_heatmaps=want_hm)
        if want_hm:
            records, heatmaps = out
            return self.pack_records(records), [PixelData(heatmaps=h) for h in heatmaps.detach()]
        return self.pack_records(out)

    def loss(self, feats, batch_data_samples, train_cfg: dict = {}):
        raise NotImplementedError("training (HeatmapHead.loss) is out of scope of the B200 inference path")
