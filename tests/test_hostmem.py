"""NUMA placement helpers for the pinned transfer buffers (gridcodegenerator_b200/hostmem.py): pure sysfs parsing,
checked against a fake /sys tree."""
import os

from gridcodegenerator_b200.hostmem import bind_to_gpu_numa_node, gpu_locality, parse_cpulist, pci_sysfs_dir


def test_parse_cpulist():
    assert parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert parse_cpulist("") == [] and parse_cpulist("5") == [5]


def test_gpu_locality_from_fake_sysfs(tmp_path):
    d = tmp_path / "0000:1b:00.0"
    d.mkdir()
    (d / "numa_node").write_text("1\n")
    (d / "local_cpulist").write_text("48-95,144-191\n")
    assert pci_sysfs_dir(0, 0x1B, 0, str(tmp_path)) == str(d)
    loc = gpu_locality(0, 0x1B, 0, str(tmp_path))
    assert loc["numa_node"] == 1 and len(loc["cpus"]) == 96 and loc["cpus"][0] == 48
    missing = gpu_locality(0, 0x2C, 0, str(tmp_path))
    assert missing["numa_node"] is None and missing["cpus"] == []
    (d / "numa_node").write_text("-1\n")               # single-node boxes report -1
    assert gpu_locality(0, 0x1B, 0, str(tmp_path))["numa_node"] is None


def test_bind_without_a_gpu_is_a_reported_no_op():
    before = os.sched_getaffinity(0)
    info = bind_to_gpu_numa_node(0)
    assert info["bound"] is False and info["why"]
    assert os.sched_getaffinity(0) == before
