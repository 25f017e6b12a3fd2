"""The console scripts the reference declares (setup.py:44-51: nhans_denoiser, nhans_separator) exist in this
package's metadata and resolve to callables with the reference's flags."""
import importlib
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_console_scripts_resolve():
    import tomllib
    with open(os.path.join(ROOT, "pyproject.toml"), "rb") as f:
        meta = tomllib.load(f)
    scripts = meta["project"]["scripts"]
    assert set(scripts) == {"nhans_denoiser", "nhans_separator"}
    for target in scripts.values():
        mod, fn = target.split(":")
        assert callable(getattr(importlib.import_module(mod), fn))
    pkgs = meta["tool"]["setuptools"]["packages"]
    for p in pkgs:
        assert os.path.isfile(os.path.join(ROOT, p.replace(".", os.sep), "__init__.py"))


def test_cli_flags_match_the_reference(capsys):
    """--input --neg --pos --output (+ --compensate --ac for the denoiser), SN/apply.py:29-35, SS/apply.py:28-34."""
    import pytest
    from nhans_b200.selective_noise import apply as sn
    from nhans_b200.source_separation import apply as ss
    for mod, flags in ((sn, ("--input", "--neg", "--pos", "--output", "--compensate", "--ac")), (ss, ("--input", "--neg", "--pos", "--output"))):
        with pytest.raises(SystemExit):
            mod.main(["--help"])
        text = capsys.readouterr().out
        for f in flags:
            assert f in text
