"""Frame kernels on the B200: bit-exact against the reference's load_video (golden checksums) and the oracle."""
import hashlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _sha(t):
    return hashlib.sha256(t.contiguous().cpu().numpy().tobytes()).hexdigest()


@pytest.mark.parametrize("name", ["448x448_m1", "896x1344_m6", "1024x1024_m6", "720x1280_m6", "1280x720_m6",
                                  "320x512_m6", "100x77_m1", "500x300_m4"])
def test_load_video_bit_exact_vs_reference_golden(built_lib, golden_frames, name):
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from mj_video_b200 import synth
    from mj_video_b200.data_processor import load_video
    g = golden_frames[name]
    v = synth.make_video(g["video_id"], n_frames=g["n_frames"], height=g["H"], width=g["W"])
    pv, counts = load_video(v, num_segments=g["num_segments"], max_num=g["max_num"])
    assert pv.device.type == "cpu" and pv.dtype == torch.float32        # reference return contract
    assert counts == g["counts"] and list(pv.shape) == g["shape"]
    assert _sha(pv) == g["sha256"]
    pv16, _ = load_video(v, num_segments=g["num_segments"], max_num=g["max_num"], device="cuda", dtype=torch.bfloat16)
    assert torch.equal(pv16.cpu(), pv.to(torch.bfloat16))               # == pixel_values.to(bfloat16) of the caller


def test_large_batch_identity_path_roundtrip(built_lib):
    """Full-size property check (256 tiles): the LUT normalisation is invertible, so frames are recovered exactly."""
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from mj_video_b200.data_processor import preprocess_frames
    g = torch.Generator(device="cuda").manual_seed(0)
    frames = torch.randint(0, 256, (256, 448, 448, 3), dtype=torch.uint8, device="cuda", generator=g)
    out, tpf = preprocess_frames(frames, 448, 1, True, torch.float32)
    assert tpf == 1 and out.shape == (256, 3, 448, 448)
    mean = torch.tensor([0.485, 0.456, 0.406], device="cuda").view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225], device="cuda").view(1, 3, 1, 1)
    back = torch.round((out * std + mean) * 255).to(torch.uint8).permute(0, 2, 3, 1)
    assert torch.equal(back, frames)
    # the reference normalises on the CPU (IEEE division; CUDA's scalar div may use a reciprocal multiply)
    ref = (frames[:4].cpu().permute(0, 3, 1, 2).float().div(255) - mean.cpu()) / std.cpu()
    assert torch.equal(out[:4].cpu(), ref)


def test_load_video_from_a_real_file(built_lib, tmp_path):
    """The reference call `load_video(path, num_segments, max_num)` (data.py:158-179) on an actual video file: host decode
    (OpenCV here), get_index sampling, then the CUDA resize / tile / normalise kernels — bit-identical to the oracle's frame
    path applied to the same decoded frames."""
    cv2 = pytest.importorskip("cv2")
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from mj_video_b200 import synth
    from mj_video_b200.data_processor import data as dp, load_video
    from oracle import port
    path = str(tmp_path / "clip.avi")
    vid = synth.make_video(6, n_frames=10, height=360, width=640)
    wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"MJPG"), 8.0, (640, 360))
    if not wr.isOpened():
        pytest.skip("this OpenCV build cannot write MJPG/AVI")
    for f in vid:
        wr.write(cv2.cvtColor(f, cv2.COLOR_RGB2BGR))
    wr.release()
    rd = dp.open_video(path)
    decoded = rd.take(range(len(rd)))
    pv, counts = load_video(path, num_segments=4, max_num=6)
    ref_pv, ref_counts = port.load_video_frames(decoded, num_segments=4, max_num=6, fps=rd.get_avg_fps())
    assert counts == ref_counts and pv.shape == ref_pv.shape and pv.dtype == torch.float32 and pv.device.type == "cpu"
    assert torch.equal(pv, ref_pv)


@pytest.mark.parametrize("H,W,max_num", [(720, 1280, 1), (720, 1280, 6), (1080, 1920, 1), (2160, 3840, 1), (77, 100, 1),
                                         (449, 447, 2), (448, 1344, 6), (1344, 448, 6), (333, 1001, 12)])
def test_fused_resize_equals_two_pass(built_lib, monkeypatch, H, W, max_num):
    """The fused resize kernel (horizontal taps into shared memory, vertical taps + crop + LUT in the same CTA) writes the
    same bits as the two-kernel path through Pillow's uint8 intermediate image, thumbnails included; odd widths exercise
    the unaligned row starts, 4K frames the smaller block heights."""
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from mj_video_b200.data_processor import preprocess_frames
    g = torch.Generator(device="cuda").manual_seed(H * 7 + W)
    frames = torch.randint(0, 256, (3, H, W, 3), dtype=torch.uint8, device="cuda", generator=g)
    monkeypatch.delenv("MJV_RESIZE_TWO_PASS", raising=False)
    fused, tpf = preprocess_frames(frames, 448, max_num, True, torch.float32)
    fused16, _ = preprocess_frames(frames, 448, max_num, True, torch.bfloat16)
    monkeypatch.setenv("MJV_RESIZE_TWO_PASS", "1")
    two, tpf2 = preprocess_frames(frames, 448, max_num, True, torch.float32)
    assert tpf == tpf2 and torch.equal(fused, two)
    assert torch.equal(fused16, two.to(torch.bfloat16))
