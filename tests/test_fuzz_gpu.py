"""GPU (-m gpu): short runs of the randomised parity drivers in profiles/ (the long runs are recorded in
profiles/r01_fuzz_parity.json).  Each driver returns 0 iff every comparison with the oracle held."""
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _gpu():
    import torch
    assert torch.cuda.is_available()
    import image_to_pointcloud_b200 as mod
    mod.load_library()


def test_hot_path_random_configurations(capsys):
    from profiles import fuzz_parity
    assert fuzz_parity.main(["fuzz_parity", "600", "4711"]) == 0
    assert fuzz_parity.main(["fuzz_parity", "12", "4712", "big"]) == 0


def test_outlier_removal_and_voxel_grid_random_clouds():
    from profiles import fuzz_rows
    assert fuzz_rows.main(["fuzz_rows", "60", "4713"]) == 0


def test_writers_random_rows():
    from profiles import fuzz_writers
    assert fuzz_writers.main(["fuzz_writers", "12", "4714"]) == 0


def test_smoothing_and_batch_api_random_frames():
    from profiles import fuzz_smooth_batch
    assert fuzz_smooth_batch.main(["fuzz_smooth_batch", "200", "4715"]) == 0
