"""TEST INFRASTRUCTURE ONLY -- stand-in for the third-party `pybullet` module.

The reference (`/root/reference/phoenix_drone_simulation`) imports PyBullet, which is
not installed in this image and is not vendored in the reference tree (setup.py:32,
un-pinned).  This module provides just enough of the PyBullet API surface for the
reference's env code to be imported *unmodified* so that golden vectors can be generated
from the reference's own arithmetic (see oracle/gen_golden.py).

Nothing under oracle/ is imported by the product package.

What is restated here (from upstream Bullet3, `examples/pybullet/pybullet.c` and
`src/LinearMath/btMatrix3x3.h`; double precision):
  * getQuaternionFromEuler   -- half-angle products, then normalisation
  * getMatrixFromQuaternion  -- btMatrix3x3::setRotation
  * getEulerFromQuaternion   -- pybullet.c, gimbal-lock branches at |sarg| >= 0.99999
Everything else is a kinematic mirror (pose/velocity setters and getters) plus, for the
`*BulletEnv` ids, a single-rigid-body integrator (`World.step`) standing in for
`stepSimulation` (semantics in SURVEY.md Appendix A.4; *parity with real Bullet is
unpinned*, there is no PyBullet binary in this image).
"""
import math
import os
import xml.etree.ElementTree as ET

import numpy as np

# --- constants touched by the reference -------------------------------------------------
GUI = 1
DIRECT = 2
COV_ENABLE_GUI = 1
COV_ENABLE_RENDERING = 7
LINK_FRAME = 1
WORLD_FRAME = 2
MAX_RAY_INTERSECTION_BATCH_SIZE = 16384
URDF_USE_INERTIA_FROM_FILE = 2
VELOCITY_CONTROL = 0
GEOM_SPHERE = 2


# --- pure helper functions -----------------------------------------------------------------
def getQuaternionFromEuler(rpy):
    roll, pitch, yaw = float(rpy[0]), float(rpy[1]), float(rpy[2])
    phi, the, psi = roll / 2.0, pitch / 2.0, yaw / 2.0
    sphi, cphi = math.sin(phi), math.cos(phi)
    sthe, cthe = math.sin(the), math.cos(the)
    spsi, cpsi = math.sin(psi), math.cos(psi)
    x = sphi * cthe * cpsi - cphi * sthe * spsi
    y = cphi * sthe * cpsi + sphi * cthe * spsi
    z = cphi * cthe * spsi - sphi * sthe * cpsi
    w = cphi * cthe * cpsi + sphi * sthe * spsi
    n = math.sqrt(x * x + y * y + z * z + w * w)
    return (x / n, y / n, z / n, w / n)


def getMatrixFromQuaternion(q):
    x, y, z, w = float(q[0]), float(q[1]), float(q[2]), float(q[3])
    d = x * x + y * y + z * z + w * w
    s = 2.0 / d
    xs, ys, zs = x * s, y * s, z * s
    wx, wy, wz = w * xs, w * ys, w * zs
    xx, xy, xz = x * xs, x * ys, x * zs
    yy, yz, zz = y * ys, y * zs, z * zs
    return (1.0 - (yy + zz), xy - wz, xz + wy,
            xy + wz, 1.0 - (xx + zz), yz - wx,
            xz - wy, yz + wx, 1.0 - (xx + yy))


def getEulerFromQuaternion(q):
    x, y, z, w = float(q[0]), float(q[1]), float(q[2]), float(q[3])
    sqx, sqy, sqz, squ = x * x, y * y, z * z, w * w
    sarg = -2.0 * (x * z - w * y)
    if sarg <= -0.99999:
        return (0.0, -0.5 * math.pi, 2.0 * math.atan2(x, -y))
    if sarg >= 0.99999:
        return (0.0, 0.5 * math.pi, 2.0 * math.atan2(-x, y))
    pitch = math.asin(sarg)
    roll = math.atan2(2.0 * (y * z + w * x), squ - sqx - sqy + sqz)
    yaw = math.atan2(2.0 * (x * y + w * z), squ + sqx - sqy - sqz)
    return (roll, pitch, yaw)


def _rot(q):
    return np.array(getMatrixFromQuaternion(q), dtype=np.float64).reshape(3, 3)


def _quat_mul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw,
        aw * bw - ax * bx - ay * by - az * bz,
    ], dtype=np.float64)


# --- a tiny kinematic / single-rigid-body world -----------------------------------------------
class Body:
    """One loaded URDF: base pose + velocity, link offsets, accumulated wrenches."""

    def __init__(self, path, pos=(0, 0, 0), quat=(0, 0, 0, 1)):
        self.path = path
        self.pos = np.array(pos, dtype=np.float64)
        self.quat = np.array(quat, dtype=np.float64)
        self.lin_vel = np.zeros(3)            # world frame
        self.ang_vel = np.zeros(3)            # world frame
        self.mass = 0.0
        self.inertia = np.zeros(3)            # principal, body frame
        self.link_offsets = []                # body-frame origin of every child link
        self.lin_damping = 0.04               # btMultiBody defaults (pybullet quickstart
        self.ang_damping = 0.04               # guide, changeDynamics: "0.04 by default")
        self.ground_z = None                  # half height of the collision cylinder
        self.clear_wrench()
        if path is not None and os.path.exists(path):
            self._parse(path)

    def _parse(self, path):
        root = ET.parse(path).getroot()
        links = {l.attrib['name']: l for l in root.findall('link')}
        base = root.findall('link')[0]
        inertial = base.find('inertial')
        if inertial is not None:
            self.mass = float(inertial.find('mass').attrib['value'])
            i = inertial.find('inertia').attrib
            self.inertia = np.array([float(i['ixx']), float(i['iyy']), float(i['izz'])])
        col = base.find('collision')
        if col is not None and col.find('geometry/cylinder') is not None:
            self.ground_z = 0.5 * float(col.find('geometry/cylinder').attrib['length'])
        for j in root.findall('joint'):
            o = j.find('origin')
            xyz = (0.0, 0.0, 0.0)
            if o is not None and 'xyz' in o.attrib:
                xyz = tuple(float(s) for s in o.attrib['xyz'].split())
            self.link_offsets.append(np.array(xyz, dtype=np.float64))
        del links

    def clear_wrench(self):
        self.force = np.zeros(3)              # world frame, at centre of mass
        self.torque = np.zeros(3)             # world frame


class World:
    def __init__(self):
        self.bodies = []
        self.dt = 1.0 / 240.0
        self.gravity = np.zeros(3)
        self.saved = {}

    def step(self):
        """Single-rigid-body stand-in for btMultiBodyDynamicsWorld::stepSimulation.

        Semi-implicit Euler in the order Bullet uses: accelerations from the wrench
        accumulated since the last step (+ gravity, + velocity damping, + gyroscopic
        term), velocities first, then position with the *new* linear velocity and the
        orientation through the exponential map of the *new* world angular velocity.
        """
        dt = self.dt
        for b in self.bodies:
            if b.mass <= 0.0:
                b.clear_wrench()
                continue
            R = _rot(b.quat)
            v_body = R.T @ b.lin_vel
            w_body = R.T @ b.ang_vel
            f_body = R.T @ (b.force + self.gravity * b.mass)
            t_body = R.T @ b.torque
            Jw = b.inertia * w_body
            # damping wrench as in btMultiBody (K1 == K2 == damping coefficient)
            f_body = f_body - b.mass * v_body * (
                b.lin_damping + b.lin_damping * np.linalg.norm(v_body))
            t_body = t_body - Jw * (
                b.ang_damping + b.ang_damping * np.linalg.norm(w_body))
            t_body = t_body - np.cross(w_body, Jw)
            # btMultiBody keeps base velocities in the world frame: the articulated-body
            # pass works in the base frame and rotates the classical accelerations back
            # (the spatial w x v term cancels), so the update is plain world-frame Euler.
            b.lin_vel = b.lin_vel + dt * (R @ (f_body / b.mass))
            b.ang_vel = b.ang_vel + dt * (R @ (t_body / b.inertia))
            b.pos = b.pos + dt * b.lin_vel
            wn = np.linalg.norm(b.ang_vel)
            if wn * dt > 1e-12:
                axis = b.ang_vel / wn
                half = 0.5 * wn * dt
                dq = np.array([*(axis * math.sin(half)), math.cos(half)])
                q = _quat_mul(dq, b.quat)
                b.quat = q / np.linalg.norm(q)
            if b.ground_z is not None and b.pos[2] < b.ground_z:
                # crude ground plane (the real engine solves a contact constraint)
                b.pos[2] = b.ground_z
                b.lin_vel[2] = max(b.lin_vel[2], 0.0)
            b.clear_wrench()


def loadURDF(fileName, basePosition=(0, 0, 0), baseOrientation=(0, 0, 0, 1), **kwargs):
    """Module-level loadURDF (base.py:211 uses it for the room): nothing to simulate."""
    return -1
