"""CPU: host-side logic of training.GraphedTrainStep that needs no device (the input refill and its checks)."""
import pytest
import torch


class _PlainBox:           # like vkit.element.Box for our purposes, but with identity equality
    def __init__(self, up, down, left, right):
        self.up, self.down, self.left, self.right = up, down, left, right


def test_graphed_step_refill_copies_tensors_and_guards_the_baked_entries():
    from vkit_ocr_model_adaptive_scaling_b200.training import GraphedTrainStep
    own = {'image': torch.zeros(2, 3, 4, 4), 'downsampled_shape': (2, 2), 'downsampled_core_box': _PlainBox(1, 2, 1, 2)}
    keep = own['image']
    new = {'image': torch.ones(2, 3, 4, 4), 'downsampled_shape': (2, 2), 'downsampled_core_box': _PlainBox(1, 2, 1, 2)}
    GraphedTrainStep._refill(own, new)                     # a fresh but equal box object with every batch is fine
    assert own['image'] is keep and bool((keep == 1).all())
    GraphedTrainStep._refill(own, own)                     # the graph's own buffers: nothing to copy
    with pytest.raises(ValueError, match='downsampled_core_box'):
        GraphedTrainStep._refill(own, {**new, 'downsampled_core_box': _PlainBox(0, 2, 1, 2)})
    with pytest.raises(ValueError, match='downsampled_shape'):
        GraphedTrainStep._refill(own, {**new, 'downsampled_shape': (4, 2)})
    with pytest.raises(ValueError, match='image'):
        GraphedTrainStep._refill(own, {**new, 'image': torch.ones(3, 3, 4, 4)})
    with pytest.raises(ValueError, match='image'):
        GraphedTrainStep._refill(own, {**new, 'image': torch.ones(2, 3, 4, 4, dtype=torch.float64)})
