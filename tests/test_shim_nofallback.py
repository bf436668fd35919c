"""skirt_b200 has no CPU fall-back (BASELINE.json north_star): a configuration outside the accelerated path, or a box without
a CUDA device, ends with the reference's fatal-error convention (SkirtCommandLineHandler.cpp:372-400: message on the console
and in the log, non-zero exit status); `--cpu` is an explicit baseline mode that says so in the log.  None of this needs a
GPU, so these tests run in the CPU suite wherever the drop-in binary has been built."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "shim", "_build", "skirt_b200")
SKI = os.path.join(ROOT, "tests", "golden", "ski")

pytestmark = pytest.mark.skipif(not os.path.exists(EXE), reason="shim/_build/skirt_b200 is built only where /root/reference exists")


def run(tmp_path, text, *flags):
    ski = tmp_path / "case.ski"
    ski.write_text(text)
    p = subprocess.run([EXE, "-t", "1", "-b", "-o", str(tmp_path), *flags, str(ski)], capture_output=True, text=True)
    log = (tmp_path / "case_log.txt").read_text() if (tmp_path / "case_log.txt").exists() else ""
    return p, log


def small(name, packets="2e4"):
    return re.sub(r'numPackets="[^"]*"', f'numPackets="{packets}"', open(os.path.join(SKI, name + ".ski")).read(), count=1)


def five_mixes(text):
    """Four copies of the GeometricMedium element with other albedos: five dust components with DIFFERENT material mixes
    (MediumSystem.cpp:874-885) -- up to SK_MAX_MEDIA = 4 run on the accelerated path (tests/golden/ski/cfg11m.ski)."""
    a, b = text.index("<GeometricMedium"), text.index("</GeometricMedium>") + len("</GeometricMedium>")
    copies = "".join(text[a:b].replace('albedos="0.6, 0.6"', 'albedos="0.%d, 0.%d"' % (k, k)) for k in (2, 3, 4, 5))
    assert 'albedos="0.6, 0.6"' in text
    return text[:b] + copies + text[b:]


@pytest.mark.parametrize("edit, reason", [
    (five_mixes, "more than 4 media with different material mixes"),
    (lambda s: s.replace('<RadiationFieldProbe', '<LaunchedPacketsProbe probeName="lpp"/><RadiationFieldProbe', 1),
     "launch call-back"),
    (lambda s: s.replace('recordPolarization="false"', 'recordPolarization="true"'), "polarization"),
])
def test_configuration_outside_the_path_is_a_fatal_error(tmp_path, edit, reason):
    text = edit(small("cfg1"))
    assert text != small("cfg1")
    p, log = run(tmp_path, text)
    assert p.returncode != 0
    assert "outside the GPU life cycle" in p.stdout and reason in p.stdout, p.stdout[-1500:]
    assert "no CPU fall-back" in log and reason in log
    assert "Finished primary emission" not in log and "GPU life cycle:" not in log
    assert not (tmp_path / "case_i60_sed.dat").exists()


def _has_gpu():
    try:
        return subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout.count("GPU ") > 0
    except OSError:
        return False


@pytest.mark.skipif(_has_gpu(), reason="needs a box WITHOUT a CUDA device")
def test_supported_configuration_without_a_gpu_fails_loudly(tmp_path):
    p, log = run(tmp_path, small("cfg1"))
    assert p.returncode != 0
    assert "no CUDA device available" in p.stdout and "no CPU fallback" in p.stdout
    assert "Finished primary emission" not in log
    assert not (tmp_path / "case_i60_sed.dat").exists()


def test_cpu_flag_is_an_explicit_baseline_mode(tmp_path):
    p, log = run(tmp_path, small("cfg1"), "--cpu")
    assert p.returncode == 0
    assert "CPU life cycle (reference)" in log and "GPU life cycle:" not in log
    assert (tmp_path / "case_i60_sed.dat").exists()


def test_duplicate_devices_are_rejected(tmp_path):
    p, _ = run(tmp_path, small("cfg1"), "-g", "0,0")
    assert p.returncode != 0 and "listed twice" in p.stdout
