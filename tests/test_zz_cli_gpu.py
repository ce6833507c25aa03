"""CLI smoke run on the GPU (`python -m tclight_b200.run --synthetic --small`): model_utils.init_synthetic ->
Generator.relight end to end on seeded random weights.  Kept in its own, last-collected file."""
import pytest

pytestmark = pytest.mark.gpu


def test_run_cli_synthetic(cuda, capsys):
    """`python -m tclight_b200.run --synthetic --small`: model_utils.init_synthetic -> Generator.relight end to end."""
    from tclight_b200 import run

    rc = run.main(["--synthetic", "--small", "--frames", "5", "--height", "176", "--width", "192", "--steps", "2", "--opt_epochs", "1"])
    out = capsys.readouterr().out
    assert rc == 0 and "finite=True" in out and "(5, 3, 176, 192)" in out
