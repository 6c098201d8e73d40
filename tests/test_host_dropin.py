"""Drop-in check of the host core (netsim_b200/host): the reference's OWN module sources
(modules/test00.c ... test05.c, compiled unchanged against netsim_b200/host/include by
netsim_b200/host/Makefile when /root/reference is present) are dlopen()ed by netsim_dev and run with
every solve / progress update / completion time / event pop on the GPU.  What a user of
`netsim modules/test0x` sees must not change: the stderr log lines and, for test05, the record file
(sha256 from SURVEY §4)."""
import hashlib
import os
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "netsim_b200", "host", "_build")
GOLD = os.path.join(ROOT, "tests", "golden")
MODS = ["test00", "test01", "test02", "test03", "test04", "test05"]


def _have():
    return os.path.exists(os.path.join(HOST, "netsim_dev")) and all(
        os.path.exists(os.path.join(HOST, "modules", m + ".so")) for m in MODS)


def test_host_core_exports_module_api(built):
    """CPU: netsim_dev exists and exports the symbols modules resolve from the executable
    (sim.h:14-23, event.h, record.h:19-22, flow.h:13, common.h:20, log.h:24)."""
    exe = os.path.join(HOST, "netsim_dev")
    if not os.path.exists(exe):
        pytest.skip("host core not built")
    syms = subprocess.run(["nm", "-D", "--defined-only", exe], capture_output=True, text=True, check=True).stdout
    for s in ("sim_create_node", "sim_register_handler", "sim_establish_flow", "sim_node_new_resource",
              "sim_node_add_resource", "sim_send_packet", "event_add", "event_remove", "flow_close", "write_record",
              "open_record", "user_done", "user_speed_change", "eps", "global_log_level"):
        assert f" {s}\n" in syms, s


@pytest.mark.gpu
@pytest.mark.parametrize("mod", MODS)
def test_unmodified_reference_module_on_device(mod, built):
    if not _have():
        pytest.skip("reference modules were not built against the host core (needs /root/reference at build time)")
    with tempfile.TemporaryDirectory(prefix="nsdrop") as d:
        p = subprocess.run([os.path.join(HOST, "netsim_dev"), os.path.join(HOST, "modules", mod)], cwd=d,
                           capture_output=True, timeout=300)
        assert p.returncode == 0, p.stderr.decode()[-500:]
        gold = open(os.path.join(GOLD, mod + ".stderr.txt"), "rb").read()
        assert p.stderr == gold
        if mod == "test05":
            data = open(os.path.join(d, "R"), "rb").read()
            assert hashlib.sha256(data).hexdigest() == "8180090a224a219634a6004be53879144c5edb2a4951c91c691039aab3135d8f"
