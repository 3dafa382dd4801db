"""TEST INFRASTRUCTURE ONLY (see mpi4py/__init__.py): a COMM_WORLD of one rank."""
import numpy as np

SUM, MIN, MAX = 'sum', 'min', 'max'


class _Comm:
    def Get_rank(self):
        return 0

    def Get_size(self):
        return 1

    def Allreduce(self, sendbuf, recvbuf, op=SUM):
        np.copyto(recvbuf, sendbuf)

    def Bcast(self, buf, root=0):
        pass

    def Gather(self, sendbuf, recvbuf, root=0):
        np.copyto(np.asarray(recvbuf).reshape(np.asarray(sendbuf).shape), sendbuf)

    def Barrier(self):
        pass


COMM_WORLD = _Comm()
