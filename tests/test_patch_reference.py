"""The reference's own classes on the B200 backend (lip2speech_b200/patch.py).  Runs where the unmodified reference is
importable (/root/reference, build container): the patched classes keep the reference's state_dict keys, route every
hot-path call into the C ABI — which refuses to run without a CUDA device, there is no CPU fallback — and unpatch()
restores the PyTorch bodies."""
import pytest
import torch

from oracle import ref_import

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="needs the reference tree (/root/reference)")


def test_patch_rebinds_reference_classes_and_fails_loudly_without_gpu():
    import lip2speech_b200.patch as b200
    from lip2speech_b200 import modules
    Decoder, VideoExtractor, SpeakerEncoder, Lip2Speech = ref_import.import_reference()
    orig = (Decoder.inference, Decoder.forward, VideoExtractor.forward, SpeakerEncoder.forward)
    done = b200.patch()
    try:
        assert {d.rsplit(".", 2)[-2] + "." + d.rsplit(".", 1)[-1] for d in done} == {
            "VideoExtractor.forward", "SpeakerEncoder.forward", "SpeakerEncoder.inference", "Decoder.inference", "Decoder.forward"}
        assert Decoder.inference is not orig[0] and VideoExtractor.forward is not orig[2]
        dec, vid = Decoder().eval(), VideoExtractor().eval()
        # same keys and shapes as the mirror modules: the backend binds weights by these names
        assert {k: tuple(v.shape) for k, v in dec.state_dict().items()} == {k: tuple(v.shape) for k, v in modules.Decoder().state_dict().items()}
        assert {k: tuple(v.shape) for k, v in vid.state_dict().items()} == {k: tuple(v.shape) for k, v in modules.VideoExtractor().state_dict().items()}
        if not torch.cuda.is_available():
            with pytest.raises(RuntimeError, match="CUDA"):
                dec.inference(torch.zeros(1, 29, 1024), torch.zeros(1, 29, 256))
            with pytest.raises(RuntimeError, match="CUDA"):
                vid(torch.zeros(1, 3, 5, 96, 96))
        if not torch.cuda.is_available():              # train mode routes into the CUDA train path as well: no CPU fallback either
            with pytest.raises(RuntimeError, match="CUDA"):
                Decoder().train()(torch.zeros(1, 29, 1024), torch.zeros(1, 29, 256), torch.zeros(1, 80, 8), None, None, 0.5)
    finally:
        b200.unpatch()
    assert (Decoder.inference, Decoder.forward, VideoExtractor.forward, SpeakerEncoder.forward) == orig
    # the restored PyTorch body runs on CPU again
    out = VideoExtractor().eval()(torch.zeros(1, 3, 2, 96, 96))
    assert out.shape == (1, 2, 768)
