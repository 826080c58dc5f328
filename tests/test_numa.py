"""Host plumbing of the frame loader: NUMA binding of a rank to its GPU's node (qcnn_gpu_b200/host/numa.py), on a fake sysfs."""
import os

from qcnn_gpu_b200.host import numa


def _fake_sysfs(tmp_path, node, cpulist):
    dev = tmp_path / "bus/pci/devices/0000:1b:00.0"
    dev.mkdir(parents=True)
    (dev / "numa_node").write_text("%d\n" % node)
    nd = tmp_path / ("devices/system/node/node%d" % max(node, 0))
    nd.mkdir(parents=True)
    (nd / "cpulist").write_text(cpulist + "\n")
    return str(tmp_path)


def test_cpulist_parsing():
    assert numa._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert numa._parse_cpulist("") == set()


def test_numa_node_lookup(tmp_path):
    sysfs = _fake_sysfs(tmp_path, 1, "0-1")
    assert numa.numa_node_of("0000:1b:00.0", sysfs) == 1
    assert numa.numa_node_of("0000:ff:00.0", sysfs) is None
    assert numa.numa_node_of(None, sysfs) is None
    assert numa.cpus_of_node(1, sysfs) == {0, 1}
    assert numa.cpus_of_node(7, sysfs) == set()


def test_negative_node_means_no_numa(tmp_path):
    sysfs = _fake_sysfs(tmp_path, -1, "0")
    assert numa.numa_node_of("0000:1b:00.0", sysfs) is None


def test_bind_without_gpu_is_a_no_op():
    before = os.sched_getaffinity(0)
    info = numa.bind_to_gpu(0)
    assert info["bound"] is False or info["node"] is not None
    if not info["bound"]:
        assert os.sched_getaffinity(0) == before
