"""Deployment surface: VehiclePredictionMsg wire format and the VehicleNode logic of
ros2_ws/src/confrez_ros (msg/VehiclePredictionMsg.msg:1-24, src/vehicle_node.py:80-189), without rclpy."""
import struct

import numpy as np
import pytest

from conflict_rez_b200.control.vehicle_node import ARRAY_FIELDS, Bool, Header, LoopbackBus, VehicleNode, VehiclePredictionMsg, deserialize, populate_msg, serialize, unpack_msg
from conflict_rez_b200.pytypes import VehiclePrediction


def test_cdr_layout_known_answer():
    """Byte layout worked out by hand from the CDR rules (alignment relative to the end of the encapsulation header)."""
    msg = VehiclePredictionMsg(header=Header(1, 2, "ab"), t=[1.0], dt=0.5, lap_num=3.0)
    raw = serialize(msg)
    exp = b"\x00\x01\x00\x00"  # CDR_LE
    exp += struct.pack("<iI", 1, 2)  # stamp
    exp += struct.pack("<I", 3) + b"ab\x00" + b"\x00"  # frame_id (length counts the NUL), pad to 4
    exp += struct.pack("<I", 1) + b"\x00" * 4 + struct.pack("<d", 1.0)  # t: count, pad to 8, data
    exp += struct.pack("<d", 0.5)  # dt
    exp += struct.pack("<I", 0) * 19  # 19 empty float64[]
    exp += b"\x00" * 4 + struct.pack("<d", 3.0)  # pad to 8, lap_num
    assert raw == exp and len(raw) == 132
    back = deserialize(raw)
    assert back == msg


def test_round_trip_and_populate_unpack():
    rng = np.random.default_rng(0)
    pred = VehiclePrediction()
    pred.x, pred.y, pred.psi = rng.normal(size=30), rng.normal(size=30), rng.normal(size=30)
    pred.dt = 0.1
    msg = populate_msg(VehiclePredictionMsg(header=Header(12, 345, "map")), pred)
    assert len(msg.x) == 30 and msg.v == [] and msg.dt == 0.1
    back = deserialize(serialize(msg))
    assert back == msg
    out = unpack_msg(back, VehiclePrediction())
    assert np.array_equal(np.array(out.x), pred.x) and np.array_equal(np.array(out.psi), pred.psi)  # float64 survives bit for bit
    assert set(ARRAY_FIELDS) <= set(VehiclePredictionMsg.__dataclass_fields__)
    with pytest.raises(ValueError):
        deserialize(b"\x00\x00\x00\x00" + serialize(msg)[4:])  # big-endian representation identifier


BACKENDS = [pytest.param("emu", id="emu"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]


@pytest.mark.parametrize("which", BACKENDS)
def test_two_vehicle_nodes_on_a_loopback_bus(request, strategy_file, which):
    """Two nodes exchange predictions through the wire format; a node steps only after every neighbour has reported in
    (vehicle_node.py:171-189), and every step solves the MPC problem with the neighbour's latest published prediction."""
    lib, dev = (request.getfixturevalue("emu_lib"), "cpu") if which == "emu" else (request.getfixturevalue("cuda_lib"), "cuda:0")
    agents = ["vehicle_1", "vehicle_2"]
    heads = {"vehicle_1": 3 * np.pi / 2, "vehicle_2": np.pi}
    bus = LoopbackBus(wire=True)
    seen = []
    bus.subscribe("/vehicle_1/pred", lambda m: seen.append(m))
    np.random.seed(0)
    nodes = [VehicleNode(bus, strategy_file, a, final_heading=heads[a], device=dev, lib=lib, agents=agents) for a in agents]
    assert len(seen) == 1 and len(seen[0].x) == 30 and bus.bytes_sent > 0
    nodes[0].timer_callback()  # vehicle_2 has not reported in yet: no step
    assert nodes[0].steps == 0
    for _ in range(4):
        for n in nodes:
            n.timer_callback()
    # vehicle_1's first tick of the loop still waits for vehicle_2's /info: 3 steps against 4
    assert nodes[0].steps == 3 and nodes[1].steps == 4
    assert len(seen) == 1 + 3
    # the neighbour's prediction each node holds is the last one published, bit for bit through the wire
    assert np.array_equal(nodes[1].vehicle.others_pred["vehicle_1"].x, np.asarray(nodes[0].vehicle.pred.x))
    assert np.array_equal(np.array(seen[-1].psi), np.asarray(nodes[0].vehicle.pred.psi))
    for n in nodes:
        assert len(n.vehicle.final_traj.x) == n.steps + 1 and n.vehicle.back_up_steps == n.vehicle.N - 1  # every solve succeeded
    assert isinstance(Bool(True).data, bool)
