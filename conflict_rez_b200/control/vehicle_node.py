"""ROS-free counterpart of the deployment surface ``ros2_ws/src/confrez_ros``.

* ``VehiclePredictionMsg`` -- the fields of ``msg/VehiclePredictionMsg.msg:1-24`` (``std_msgs/Header`` + 20 ``float64[]`` + ``dt`` +
  ``lap_num``), with ``populate_msg`` / ``unpack_msg`` (what ``MPClabNode`` does for same-named attributes, ``vehicle_node.py:143-150,
  154-163``) and the message's CDR wire encoding (``serialize`` / ``deserialize``: little-endian XCDR1 as rmw puts it on the wire --
  4-byte encapsulation header, every primitive aligned to its size relative to the end of that header, sequences as uint32 count +
  elements, strings as uint32 length including the terminating NUL).
* ``VehicleNode`` -- the logic of ``src/vehicle_node.py:80-189`` over an abstract bus (``publish(topic, msg)`` /
  ``subscribe(topic, callback)``): publishes ``/<agent>/pred`` and ``/<agent>/info``, listens to the other vehicles, and runs one MPC
  step per timer tick once every neighbour has reported in.  ``LoopbackBus`` delivers in-process (optionally through the wire
  encoding); a ROS 2 deployment passes a thin wrapper around ``rclpy`` publishers/subscriptions instead (rclpy is not in this image).

The MPC solve inside ``VehicleFollower.step`` is the CUDA kernel (``ObcaMpcSolver``); nothing here touches the hot path.
"""
import struct
from dataclasses import dataclass, field
from typing import Callable, Dict, List

import numpy as np

from conflict_rez_b200.control.vehicle_follower import VehicleFollower
from conflict_rez_b200.pytypes import VehiclePrediction, VehicleState

# float64[] fields in declaration order (VehiclePredictionMsg.msg:3-23); `dt` sits after `t`, `lap_num` is last
ARRAY_FIELDS = ["t", "x", "y", "v", "v_x", "v_y", "a_y", "a_x", "psi", "psidot", "s", "x_tran", "v_long", "v_tran", "a_long", "a_tran", "e_psi",
                "u_a", "u_steer", "u_steer_dot"]


@dataclass
class Header:  # std_msgs/Header
    sec: int = 0
    nanosec: int = 0
    frame_id: str = ""


@dataclass
class VehiclePredictionMsg:
    header: Header = field(default_factory=Header)
    t: List[float] = field(default_factory=list)
    dt: float = 0.0
    x: List[float] = field(default_factory=list)
    y: List[float] = field(default_factory=list)
    v: List[float] = field(default_factory=list)
    v_x: List[float] = field(default_factory=list)
    v_y: List[float] = field(default_factory=list)
    a_y: List[float] = field(default_factory=list)
    a_x: List[float] = field(default_factory=list)
    psi: List[float] = field(default_factory=list)
    psidot: List[float] = field(default_factory=list)
    s: List[float] = field(default_factory=list)
    x_tran: List[float] = field(default_factory=list)
    v_long: List[float] = field(default_factory=list)
    v_tran: List[float] = field(default_factory=list)
    a_long: List[float] = field(default_factory=list)
    a_tran: List[float] = field(default_factory=list)
    e_psi: List[float] = field(default_factory=list)
    u_a: List[float] = field(default_factory=list)
    u_steer: List[float] = field(default_factory=list)
    u_steer_dot: List[float] = field(default_factory=list)
    lap_num: float = 0.0


@dataclass
class Bool:  # std_msgs/Bool
    data: bool = False


def populate_msg(msg: VehiclePredictionMsg, pred: VehiclePrediction) -> VehiclePredictionMsg:
    """Copy every attribute the message and the prediction share and that is set (MPClabNode.populate_msg semantics)."""
    for k in ARRAY_FIELDS:
        v = getattr(pred, k, None)
        if v is not None:
            setattr(msg, k, [float(a) for a in np.ravel(v)])
    for k in ("dt", "lap_num"):
        v = getattr(pred, k, None)
        if v is not None:
            setattr(msg, k, float(v))
    return msg


def unpack_msg(msg: VehiclePredictionMsg, pred: VehiclePrediction) -> VehiclePrediction:
    for k in ARRAY_FIELDS:
        if hasattr(pred, k):
            setattr(pred, k, list(getattr(msg, k)))
    for k in ("dt", "lap_num"):
        if hasattr(pred, k):
            setattr(pred, k, getattr(msg, k))
    return pred


class _Cdr:
    """Little-endian CDR stream; alignment is relative to the first byte after the 4-byte encapsulation header."""

    def __init__(self, data: bytes = b""):
        self.buf = bytearray(data)
        self.pos = 0

    def _pad(self, n):
        r = len(self.buf) % n
        if r:
            self.buf.extend(b"\x00" * (n - r))

    def put(self, fmt, size, *vals):
        self._pad(size)
        self.buf.extend(struct.pack("<" + fmt, *vals))

    def put_f64_seq(self, vals):
        self.put("I", 4, len(vals))
        if len(vals):
            self._pad(8)
            self.buf.extend(np.asarray(vals, dtype="<f8").tobytes())

    def get(self, fmt, size):
        self.pos += (-self.pos) % size
        (v,) = struct.unpack_from("<" + fmt, self.buf, self.pos)
        self.pos += size
        return v

    def get_f64_seq(self):
        n = self.get("I", 4)
        if n == 0:
            return []
        self.pos += (-self.pos) % 8
        out = np.frombuffer(bytes(self.buf[self.pos : self.pos + 8 * n]), dtype="<f8").tolist()
        self.pos += 8 * n
        return out


CDR_LE = b"\x00\x01\x00\x00"  # representation identifier CDR_LE + options


def serialize(msg: VehiclePredictionMsg) -> bytes:
    c = _Cdr()
    c.put("i", 4, msg.header.sec)
    c.put("I", 4, msg.header.nanosec)
    fid = msg.header.frame_id.encode() + b"\x00"
    c.put("I", 4, len(fid))
    c.buf.extend(fid)
    c.put_f64_seq(msg.t)
    c.put("d", 8, msg.dt)
    for k in ARRAY_FIELDS[1:]:
        c.put_f64_seq(getattr(msg, k))
    c.put("d", 8, msg.lap_num)
    return CDR_LE + bytes(c.buf)


def deserialize(data: bytes) -> VehiclePredictionMsg:
    if data[:2] != CDR_LE[:2]:
        raise ValueError("VehiclePredictionMsg: only little-endian CDR is supported")
    c = _Cdr(data[4:])
    msg = VehiclePredictionMsg()
    msg.header.sec = c.get("i", 4)
    msg.header.nanosec = c.get("I", 4)
    n = c.get("I", 4)
    msg.header.frame_id = bytes(c.buf[c.pos : c.pos + n - 1]).decode()
    c.pos += n
    msg.t = c.get_f64_seq()
    msg.dt = c.get("d", 8)
    for k in ARRAY_FIELDS[1:]:
        setattr(msg, k, c.get_f64_seq())
    msg.lap_num = c.get("d", 8)
    return msg


class LoopbackBus:
    """In-process topic bus.  ``wire=True`` sends every VehiclePredictionMsg through serialize/deserialize like a real transport."""

    def __init__(self, wire: bool = False):
        self.subs: Dict[str, List[Callable]] = {}
        self.last: Dict[str, object] = {}  # latched: a late subscriber gets the last message of its topic (nodes are built one by one here)
        self.wire, self.bytes_sent = wire, 0

    def subscribe(self, topic: str, callback: Callable):
        self.subs.setdefault(topic, []).append(callback)
        if topic in self.last:
            callback(self.last[topic])

    def publish(self, topic: str, msg):
        if self.wire and isinstance(msg, VehiclePredictionMsg):
            raw = serialize(msg)
            self.bytes_sent += len(raw)
            msg = deserialize(raw)
        self.last[topic] = msg
        for cb in self.subs.get(topic, []):
            cb(msg)


class VehicleNode:
    """One path-following vehicle on a bus (``vehicle_node.py:80-189``): topics ``/<agent>/pred`` (VehiclePredictionMsg) and
    ``/<agent>/info`` (Bool); ``timer_callback`` is what the 0.05 s ROS timer runs."""

    def __init__(self, bus, rl_file_name: str, agent: str, num_vehicles: int = 4, final_heading: float = None, spline_ws: bool = True,
                 init_offset: VehicleState = None, device="cuda:0", lib=None, color=None, agents: List[str] = None):
        self.bus, self.agent = bus, agent
        self.vehicle = VehicleFollower(rl_file_name, agent=agent, color=color or {}, init_offset=init_offset or VehicleState(), final_heading=final_heading,
                                       device=device)
        if lib is not None:
            self.vehicle._lib = lib
        names = agents if agents is not None else ["vehicle_%d" % i for i in range(num_vehicles)]  # vehicle_node.py:117-121
        self.others = [a for a in names if a != agent]
        self.others_info = {o: False for o in self.others}
        for other in self.others:
            bus.subscribe("/%s/pred" % other, self.vehicle_pred_cb(other))
            bus.subscribe("/%s/info" % other, self.vehicle_info_cb(other))
        self.vehicle.others = self.others
        self.vehicle.plan_single_path(spline_ws=spline_ws)
        self.vehicle.setup_controller()
        self.vehicle.get_current_ref()
        self.steps = 0
        self.publish_pred()

    def publish_pred(self):
        pred = VehiclePrediction()
        pred.x, pred.y, pred.psi = self.vehicle.pred.x, self.vehicle.pred.y, self.vehicle.pred.psi
        self.bus.publish("/%s/pred" % self.agent, populate_msg(VehiclePredictionMsg(), pred))

    def vehicle_pred_cb(self, other):
        def callback(msg):
            pred = VehiclePrediction()
            unpack_msg(msg, pred)
            pred.x, pred.y, pred.psi = np.array(pred.x), np.array(pred.y), np.array(pred.psi)
            self.vehicle.others_pred[other] = pred

        return callback

    def vehicle_info_cb(self, other):
        def callback(msg):
            self.others_info[other] = bool(msg.data)

        return callback

    def timer_callback(self):
        self.bus.publish("/%s/info" % self.agent, Bool(True))
        if all(self.others_info.values()) and all(o in self.vehicle.others_pred for o in self.others):
            self.vehicle.step()
            self.steps += 1
            self.publish_pred()
