"""Native whole-frame DiMP tracker behind the reference's tracker plug-in interface (`BaseTracker`:
pytracking/tracker/base/basetracker.py:6-22 -- `initialize(image, info) -> dict`, `track(image, info) -> {'target_bbox': [x,y,w,h]}`).

One `track()` = one C-ABI call (`b200trk_dimp_track_host`, include/b200trk.h): the uint8 frame goes to the GPU, the crop is sampled
there, backbone + head + classify + localisation run as kernels, 64 bytes come back, the float32 tracker state is advanced on the
host inside the library, and the online filter update is left running on the stream.  Nothing here falls back to PyTorch."""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from .frame_engine import DiMPFrameEngine

FLAGS = {0: None, 1: "normal", 2: "hard_negative", 3: "uncertain", 4: "not_found"}

_DEFAULTS = dict(                                    # pytracking/parameter/dimp/dimp50.py + the `.get` defaults of dimp.py
    image_sample_size=288, search_area_scale=5, sample_memory_size=50, learning_rate=0.01, hard_negative_learning_rate=None,
    init_samples_minimum_weight=None, train_skipping=1, train_sample_interval=1, net_opt_iter=None, net_opt_update_iter=None,
    net_opt_hn_iter=None, update_classifier=False, advanced_localization=False, target_not_found_threshold=0.0,
    distractor_threshold=0.0, hard_negative_threshold=0.0, target_neighborhood_scale=0.0, dispalcement_scale=0.0,
    uncertain_threshold=-math.inf, hard_sample_threshold=-math.inf, target_inside_ratio=0.2, augmentation_expansion_factor=None,
    output_not_found_box=False, use_iou_net=True, iounet_k=5, num_init_random_boxes=0, box_jitter_pos=0.0, box_jitter_sz=0.0,
    maximal_aspect_ratio=6.0, box_refinement_iter=0, box_refinement_step_length=1.0, box_refinement_step_decay=1.0,
    box_refinement_space="default", update_scale_when_uncertain=True, use_iounet_pos_for_learning=True)


def make_params(source=None, **overrides):
    """b200trk_dimp_params_t from a reference `TrackerParams` object (attribute access) and / or keyword overrides."""
    vals = dict(_DEFAULTS)
    if source is not None:
        for k in vals:
            if hasattr(source, k):
                vals[k] = getattr(source, k)
    vals.update(overrides)
    p = _lib.DimpParams()
    none_int = lambda v: -1 if v is None else int(v)
    p.image_sample_size = int(vals["image_sample_size"])
    p.search_area_scale = float(vals["search_area_scale"])
    p.sample_memory_size = int(vals["sample_memory_size"])
    p.learning_rate = float(vals["learning_rate"])
    p.hard_negative_learning_rate = -1.0 if vals["hard_negative_learning_rate"] is None else float(vals["hard_negative_learning_rate"])
    p.init_samples_minimum_weight = float(vals["init_samples_minimum_weight"] or 0.0)
    p.train_skipping, p.train_sample_interval = int(vals["train_skipping"]), int(vals["train_sample_interval"])
    p.net_opt_iter, p.net_opt_update_iter = none_int(vals["net_opt_iter"]), none_int(vals["net_opt_update_iter"])
    p.net_opt_hn_iter = none_int(vals["net_opt_hn_iter"])
    p.update_classifier, p.advanced_localization = int(bool(vals["update_classifier"])), int(bool(vals["advanced_localization"]))
    for k in ("target_not_found_threshold", "distractor_threshold", "hard_negative_threshold", "target_neighborhood_scale",
              "dispalcement_scale", "uncertain_threshold", "hard_sample_threshold", "target_inside_ratio"):
        setattr(p, k, float(vals[k]))
    p.augmentation_expansion_factor = float(vals["augmentation_expansion_factor"] or 0.0)
    p.output_not_found_box = int(bool(vals["output_not_found_box"]))
    p.use_iou_net, p.iounet_k = int(bool(vals["use_iou_net"])), int(vals["iounet_k"])
    p.num_init_random_boxes = int(vals["num_init_random_boxes"])
    p.box_jitter_pos, p.box_jitter_sz = float(vals["box_jitter_pos"]), float(vals["box_jitter_sz"])
    p.maximal_aspect_ratio = float(vals["maximal_aspect_ratio"])
    p.box_refinement_iter = int(vals["box_refinement_iter"])
    if isinstance(vals["box_refinement_step_length"], (tuple, list)):
        raise NotImplementedError("b200trk: per-coordinate box_refinement_step_length")
    p.box_refinement_step_length, p.box_refinement_step_decay = float(vals["box_refinement_step_length"]), float(vals["box_refinement_step_decay"])
    p.box_refinement_relative = int(vals["box_refinement_space"] == "relative")
    p.update_scale_when_uncertain = int(bool(vals["update_scale_when_uncertain"]))
    p.use_iounet_pos_for_learning = int(bool(vals["use_iounet_pos_for_learning"]))
    return p


class HostLogic:
    """The host half of the tracker alone (no GPU): crop planning and the post-localisation state update.  Used by the CPU tests."""

    def __init__(self, params):
        self.params = params
        h = C.c_void_p()
        _lib.check(_lib.lib().b200trk_dimp_tracker_create(C.byref(h), None, C.byref(params)), "dimp_tracker_create")
        self.handle = h

    def init_state(self, H, W, init_bbox):
        g, box = _lib.CropGeom(), (C.c_float * 4)()
        bb = (C.c_double * 4)(*[float(v) for v in init_bbox])
        _lib.check(_lib.lib().b200trk_dimp_tracker_init_state(self.handle, H, W, C.byref(bb), C.byref(g), C.byref(box)), "init_state")
        return g, np.array(box, dtype=np.float32)

    def adopt(self, H, W, state9, sample_weights, num_stored, num_init, previous_replace_ind=-1, frame_num=1):
        s = [float(v) for v in state9]
        sw = np.ascontiguousarray(sample_weights, dtype=np.float32)
        _lib.check(_lib.lib().b200trk_dimp_tracker_adopt(
            self.handle, H, W, C.byref((C.c_float * 2)(s[0], s[1])), C.byref((C.c_float * 2)(s[2], s[3])), s[4],
            C.byref((C.c_float * 2)(s[5], s[6])), s[7], s[8], sw.ctypes.data_as(C.c_void_p), int(num_stored), int(num_init),
            int(previous_replace_ind), int(frame_num)), "adopt")

    def plan_crop(self):
        g = _lib.CropGeom()
        _lib.check(_lib.lib().b200trk_dimp_tracker_plan_crop(self.handle, C.byref(g)), "plan_crop")
        return g

    def commit(self, geom, loc):
        info = _lib.FrameInfo()
        sw = np.zeros(self.params.sample_memory_size, dtype=np.float32)
        _lib.check(_lib.lib().b200trk_dimp_tracker_commit(self.handle, C.byref(geom), C.byref(loc), C.byref(info),
                                                          sw.ctypes.data_as(C.c_void_p)), "commit")
        return info, sw

    def state(self):
        out = (C.c_float * 9)()
        _lib.check(_lib.lib().b200trk_dimp_tracker_state(self.handle, C.byref(out)), "state")
        return np.array(out, dtype=np.float32)

    def close(self):
        if getattr(self, "handle", None):
            _lib.lib().b200trk_dimp_tracker_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DiMPTracker(HostLogic):
    """initialize / track on one GPU.  `state_dict`: the DiMP network's state_dict (reference key names)."""

    def __init__(self, state_dict, params=None, arch="resnet50", precision=0, device=None, **param_overrides):
        self.params = params if isinstance(params, _lib.DimpParams) else make_params(params, **param_overrides)
        p = self.params
        self.engine = DiMPFrameEngine(state_dict, arch=arch, filter_size=4, memory_size=p.sample_memory_size, max_batch=1,
                                      crop_size=p.image_sample_size, precision=precision, device=device)
        self.device = self.engine.device
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().b200trk_dimp_tracker_create(C.byref(h), self.engine.state, C.byref(p)), "dimp_tracker_create")
        self.handle = h
        self.info = _lib.FrameInfo()
        self.debug_info = {}
        self._pinned = None
        self.iou = None
        self.torch_noise = False       # True: draw the random proposals' noise with torch.rand, exactly where the reference does

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _image(self, image):
        if isinstance(image, torch.Tensor):                     # a pinned uint8 [H,W,3] tensor is used in place
            if image.dtype != torch.uint8 or image.dim() != 3 or image.shape[2] != 3 or image.is_cuda or not image.is_contiguous():
                raise RuntimeError("DiMPTracker: image must be a contiguous uint8 [H,W,3] host tensor or ndarray")
            return image, image.data_ptr(), image.shape[0], image.shape[1]
        a = np.ascontiguousarray(image)
        if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != 3:
            raise RuntimeError("DiMPTracker: image must be uint8 [H,W,3] (RGB), got %s %s" % (a.dtype, a.shape))
        if self._pinned is None or tuple(self._pinned.shape) != a.shape:
            self._pinned = torch.empty(a.shape, dtype=torch.uint8).pin_memory()
        self._pinned.numpy()[...] = a                            # pageable -> pinned staging, then one asynchronous H2D in the library
        return self._pinned, self._pinned.data_ptr(), a.shape[0], a.shape[1]

    def initialize(self, image, info):
        """DiMP.initialize for the un-augmented configuration (see b200trk_dimp_tracker_initialize_host)."""
        keep, ptr, H, W = self._image(image)
        bb = (C.c_double * 4)(*[float(v) for v in info["init_bbox"]])
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().b200trk_dimp_tracker_initialize_host(self.handle, C.c_void_p(ptr), H, W, C.byref(bb), self._stream()),
                       "dimp_tracker_initialize_host")
        return {}

    def attach_iounet(self, state_dict, modulation):
        """IoUNet refinement: `state_dict` of the whole DiMP network (bb_regressor.* keys), `modulation` = the two modulation vectors
        of the first-frame target (AtomIoUNet.get_modulation; DiMP.init_iou_net dimp.py:509-540)."""
        from .iou import IoUPredictor
        if not getattr(self.engine.backbone, "iou_dims", None):
            self.engine.backbone.attach_iou_head(state_dict)
        self.iou = IoUPredictor(state_dict, device=self.device)
        m3 = modulation[0].detach().float().reshape(-1).cpu().contiguous()
        m4 = modulation[1].detach().float().reshape(-1).cpu().contiguous()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().b200trk_dimp_tracker_attach_iounet(self.handle, self.iou.handle, C.c_void_p(m3.data_ptr()),
                                                                     C.c_void_p(m4.data_ptr())), "dimp_tracker_attach_iounet")

    def adopt_reference(self, ref, image_hw):
        """Take over from a reference `DiMP` object after its `initialize()` (any augmentation / initialiser): scalars, sample
        weights, sample memory, boxes and filter are copied into the engine; tracking continues natively."""
        e = self.engine
        n = int(min(int(ref.num_stored_samples[0]), self.params.sample_memory_size))
        e.memory[:n].copy_(ref.training_samples[0][:n].to(self.device))
        e.boxes[:n].copy_(ref.target_boxes[:n].to(self.device))
        e.sample_weights.copy_(ref.sample_weights[0].to(self.device))
        e.filter.copy_(ref.target_filter.reshape(e.filter.shape).to(self.device))
        st = [float(ref.pos[0]), float(ref.pos[1]), float(ref.target_sz[0]), float(ref.target_sz[1]), float(ref.target_scale),
              float(ref.base_target_sz[0]), float(ref.base_target_sz[1]), float(ref.min_scale_factor), float(ref.max_scale_factor)]
        prev = ref.previous_replace_ind[0]
        self.adopt(int(image_hw[0]), int(image_hw[1]), st, ref.sample_weights[0].detach().float().cpu().numpy(),
                   int(ref.num_stored_samples[0]), int(ref.num_init_samples[0]), -1 if prev is None else int(prev), int(ref.frame_num))
        if self.params.use_iou_net and self.iou is None:
            self.attach_iounet(ref.net.net.state_dict(), ref.iou_modulation)
        torch.cuda.synchronize(self.device)

    def _noise(self):
        n = self.params.num_init_random_boxes
        if self.torch_noise and self.params.use_iou_net and n > 0:
            u = torch.rand(n, 4).contiguous()                          # dimp.py:667
            _lib.check(_lib.lib().b200trk_dimp_tracker_set_proposal_noise(self.handle, C.c_void_p(u.data_ptr()), 4 * n), "set_proposal_noise")

    def track(self, image, info=None):
        self._noise()
        keep, ptr, H, W = self._image(image)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().b200trk_dimp_track_host(self.handle, C.c_void_p(ptr), H, W, C.byref(self.info), self._stream()),
                       "dimp_track_host")
        self.debug_info = {"flag": FLAGS[self.info.flag], "max_score": float(self.info.max_score),
                           "predicted_iou": float(self.info.predicted_iou) if self.info.refined else None}
        return {"target_bbox": [float(v) for v in self.info.bbox]}

    def track_device(self, image_dev):
        """`track` with the uint8 [H,W,3] frame already on the GPU (a CUDA tensor)."""
        self._noise()
        if not image_dev.is_cuda or image_dev.dtype != torch.uint8 or image_dev.dim() != 3 or not image_dev.is_contiguous():
            raise RuntimeError("DiMPTracker.track_device: contiguous uint8 [H,W,3] CUDA tensor")
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().b200trk_dimp_track_device(self.handle, C.c_void_p(image_dev.data_ptr()), image_dev.shape[0],
                                                            image_dev.shape[1], C.byref(self.info), self._stream()), "dimp_track_device")
        return self.info

    def close(self):
        HostLogic.close(self)
        if getattr(self, "iou", None) is not None:
            self.iou.close()
            self.iou = None
        if getattr(self, "engine", None) is not None:
            self.engine.close()
            self.engine = None
