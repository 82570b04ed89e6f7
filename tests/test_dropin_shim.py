"""The reference driver, UNMODIFIED, on top of the drop-in shim.

/root/reference exists only in the authoring container, which has no GPU; the GPU box has no reference.  So the
wiring is checked here: the real describe_clip_neurons.py (stub `clip` / `data_utils`, as SURVEY.md section 4
describes) must resolve `similarity.*`, `CLIP_og_utils.get_activation` and
`CLIP_og_utils.get_similarity_from_activations` to the B200 implementations -- which then refuse to run without
CUDA.  The numerical end-to-end check of the same call sequence runs on the GPU (test_gpu_parity.py::
test_driver_call_sequence)."""
import os
import subprocess
import sys
import textwrap

import pytest

REF = "/root/reference/concept_vit"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "describe_clip_neurons.py")),
                                reason="reference tree not present (GPU box)")

STUB_CLIP = """
import torch
class _M(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.proj = torch.nn.Linear(3 * 8 * 8, 16)
    def encode_image(self, x):
        return self.proj(x.flatten(1))
    def encode_text(self, t):
        return torch.nn.functional.one_hot(t[:, 0] % 16, 16).float() + 0.1
def load(name, device="cpu"):
    return _M().to(device).eval(), None
def tokenize(words):
    return torch.tensor([[sum(map(ord, w)) % 97, 0] for w in words])
"""
STUB_DATA = """
import torch
class _T(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.layer1 = torch.nn.Conv2d(3, 6, 3)
        self.fc = torch.nn.Linear(6 * 6 * 6, 5)
    def forward(self, x):
        return self.fc(self.layer1(x).flatten(1))
def get_target_model(name, device):
    return _T().to(device).eval(), None
def get_data(name, preprocess=None):
    g = torch.Generator().manual_seed(0)
    return [(torch.randn(3, 8, 8, generator=g), 0) for _ in range(120)]
"""


def _run(tmp_path, device):
    stubs = tmp_path / "stubs"
    stubs.mkdir()
    (stubs / "clip.py").write_text(textwrap.dedent(STUB_CLIP))
    (stubs / "data_utils.py").write_text(textwrap.dedent(STUB_DATA))
    (tmp_path / "concepts.txt").write_text("\n".join("concept%d" % i for i in range(11)))
    env = dict(os.environ, MCD_EXTRA_PATH=str(stubs), PYTHONPATH=ROOT)
    cmd = [sys.executable, "-m", "mammo_clip_dissect_b200.shim.run_reference_driver",
           os.path.join(REF, "describe_clip_neurons.py"), "--target_layers", "layer1,fc", "--d_probe", "broden",
           "--concept_set", str(tmp_path / "concepts.txt"), "--device", device, "--batch_size", "50",
           "--activation_dir", str(tmp_path / "act"), "--result_dir", str(tmp_path / "res")]
    return subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)


def test_unmodified_driver_resolves_to_the_b200_path(tmp_path):
    r = _run(tmp_path, "cpu")
    assert r.returncode != 0
    # the forward hook registered by the reference's save_target_activations is ours ...
    assert "mammo_clip_dissect_b200 has no CPU path" in r.stderr, r.stderr[-2000:]
    assert "hooks.py" in r.stderr or "similarity.py" in r.stderr or "features.py" in r.stderr
