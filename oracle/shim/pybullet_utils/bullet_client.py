"""TEST INFRASTRUCTURE ONLY -- stand-in for `pybullet_utils.bullet_client`.

One `BulletClient` owns one `pybullet.World`.  Methods the reference calls
(SURVEY.md section 8c lists the measured API surface) are implemented as a kinematic
mirror; force/torque application and `stepSimulation` drive the single-rigid-body
integrator in oracle/shim/pybullet.py (used by the `*BulletEnv` ids only).
"""
import copy
import os

import numpy as np
import pybullet as pb


class BulletClient:
    GEOM_SPHERE = pb.GEOM_SPHERE
    COV_ENABLE_RENDERING = pb.COV_ENABLE_RENDERING
    COV_ENABLE_GUI = pb.COV_ENABLE_GUI

    def __init__(self, connection_mode=None):
        self.world = pb.World()
        self.search_paths = []
        self._n_states = 0

    # -- configuration ------------------------------------------------------------------
    def setAdditionalSearchPath(self, path):
        self.search_paths.append(path)

    def configureDebugVisualizer(self, *a, **k):
        pass

    def resetDebugVisualizerCamera(self, *a, **k):
        pass

    def setPhysicsEngineParameter(self, fixedTimeStep=None, **k):
        if fixedTimeStep is not None:
            self.world.dt = float(fixedTimeStep)

    def setGravity(self, x, y, z):
        self.world.gravity = np.array([x, y, z], dtype=np.float64)

    def disconnect(self):
        pass

    # -- bodies -----------------------------------------------------------------------------
    def _resolve(self, name):
        if os.path.isabs(name) and os.path.exists(name):
            return name
        for p in reversed(self.search_paths):
            cand = os.path.join(p, name)
            if os.path.exists(cand):
                return cand
        return None

    def loadURDF(self, fileName, basePosition=(0, 0, 0), baseOrientation=(0, 0, 0, 1),
                 **kwargs):
        body = pb.Body(self._resolve(fileName), basePosition, baseOrientation)
        if os.path.basename(fileName) == 'plane.urdf':
            body.mass = 0.0
        self.world.bodies.append(body)
        return len(self.world.bodies) - 1

    def createVisualShape(self, *a, **k):
        return -1

    def createMultiBody(self, baseMass=0, basePosition=(0, 0, 0), **k):
        body = pb.Body(None, basePosition)
        self.world.bodies.append(body)
        return len(self.world.bodies) - 1

    def changeDynamics(self, bodyUniqueId, linkIndex, mass=None,
                       localInertiaDiagonal=None, **k):
        b = self.world.bodies[bodyUniqueId]
        if linkIndex == -1:
            if mass is not None:
                b.mass = float(mass)
            if localInertiaDiagonal is not None:
                b.inertia = np.array(localInertiaDiagonal, dtype=np.float64)

    # -- state save / restore -----------------------------------------------------------
    def saveState(self):
        self._n_states += 1
        self.world.saved[self._n_states] = copy.deepcopy(
            [(b.pos, b.quat, b.lin_vel, b.ang_vel) for b in self.world.bodies])
        return self._n_states

    def restoreState(self, stateId):
        snap = self.world.saved[stateId]
        for b, (p, q, v, w) in zip(self.world.bodies, snap):
            b.pos, b.quat = p.copy(), q.copy()
            b.lin_vel, b.ang_vel = v.copy(), w.copy()
            b.clear_wrench()

    # -- kinematic mirror -----------------------------------------------------------------
    def resetBasePositionAndOrientation(self, bodyUniqueId, posObj, ornObj):
        b = self.world.bodies[bodyUniqueId]
        b.pos = np.array(posObj, dtype=np.float64)
        b.quat = np.array(ornObj, dtype=np.float64)

    def resetBaseVelocity(self, objectUniqueId, linearVelocity=None, angularVelocity=None):
        b = self.world.bodies[objectUniqueId]
        if linearVelocity is not None:
            b.lin_vel = np.array(linearVelocity, dtype=np.float64)
        if angularVelocity is not None:
            b.ang_vel = np.array(angularVelocity, dtype=np.float64)

    def getBasePositionAndOrientation(self, bodyUniqueId):
        b = self.world.bodies[bodyUniqueId]
        return tuple(b.pos.tolist()), tuple(b.quat.tolist())

    def getBaseVelocity(self, bodyUniqueId):
        b = self.world.bodies[bodyUniqueId]
        return tuple(b.lin_vel.tolist()), tuple(b.ang_vel.tolist())

    def getLinkStates(self, bodyUniqueId, linkIndices, **k):
        b = self.world.bodies[bodyUniqueId]
        R = pb._rot(b.quat)
        out = []
        for i in linkIndices:
            p = b.pos + R @ b.link_offsets[i]
            out.append((tuple(p.tolist()), tuple(b.quat.tolist())))
        return out

    # -- pure helpers ---------------------------------------------------------------------
    getQuaternionFromEuler = staticmethod(pb.getQuaternionFromEuler)
    getMatrixFromQuaternion = staticmethod(pb.getMatrixFromQuaternion)
    getEulerFromQuaternion = staticmethod(pb.getEulerFromQuaternion)

    # -- dynamics (Bullet ids) --------------------------------------------------------------
    def applyExternalForce(self, objectUniqueId, linkIndex, forceObj, posObj, flags):
        b = self.world.bodies[objectUniqueId]
        R = pb._rot(b.quat)
        off = b.link_offsets[linkIndex] if linkIndex >= 0 else np.zeros(3)
        f = np.array(forceObj, dtype=np.float64)
        p = np.array(posObj, dtype=np.float64)
        if flags == pb.LINK_FRAME:
            f_world = R @ f
            r_world = R @ (off + p)        # all child links share the base orientation
        else:
            f_world = f
            r_world = p - b.pos
        b.force = b.force + f_world
        b.torque = b.torque + np.cross(r_world, f_world)

    def applyExternalTorque(self, objectUniqueId, linkIndex, torqueObj, flags):
        b = self.world.bodies[objectUniqueId]
        t = np.array(torqueObj, dtype=np.float64)
        if flags == pb.LINK_FRAME:
            t = pb._rot(b.quat) @ t
        b.torque = b.torque + t

    def setJointMotorControl2(self, *a, **k):
        pass   # propeller spin is visual only (link inertia 1e-9), see SURVEY A.4

    def stepSimulation(self):
        self.world.step()

    def addUserDebugLine(self, *a, **k):
        return -1

    def removeAllUserDebugItems(self):
        pass
