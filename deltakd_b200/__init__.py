"""deltakd_b200 — B200-native (sm_100a) implementation of DeltaKD's distillation-loss hot path.

Public surface mirrors /root/reference/model/{loss,misc,models}.py for that path:
`DistillationLoss`, `call_base_loss`, the per-method loss functions, `random_masking`,
`saliency_masking`, `forward_with_features` and the KD-head attachment.  Importing
`deltakd_b200.loss` loads libdeltakd_sm100.so and fails loudly if it has not been built.
"""
__all__ = ["DistillationLoss", "call_base_loss", "random_masking", "saliency_masking",
           "forward_with_features", "attach_distillation_heads", "FrozenTeacher", "needed_layers",
           "Mixup", "MixedLabels", "FusedStepEpilogue", "accuracy"]


def __getattr__(name):  # lazy: `import deltakd_b200.synth` must work without the CUDA library
    if name in ("DistillationLoss", "call_base_loss", "lrkd_loss", "curkd_loss", "mgd_loss",
                "saliency_mgd_loss", "vitkd_loss", "diffkd_loss", "SoftTargetCrossEntropy", "LabelSmoothingCrossEntropy"):
        from . import loss
        return getattr(loss, name)
    if name in ("random_masking", "saliency_masking"):
        from . import misc
        return getattr(misc, name)
    if name == "forward_with_features":
        from .features import forward_with_features
        return forward_with_features
    if name in ("FrozenTeacher", "needed_layers"):
        from . import features
        return getattr(features, name)
    if name in ("Mixup", "MixedLabels"):
        from . import mixup
        return getattr(mixup, name)
    if name in ("FusedStepEpilogue", "accuracy"):
        from . import step
        return getattr(step, name)
    if name == "attach_distillation_heads":
        from .heads import attach_distillation_heads
        return attach_distillation_heads
    raise AttributeError(name)
