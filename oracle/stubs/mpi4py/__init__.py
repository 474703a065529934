"""Single-process stand-in for mpi4py (absent from the image): what the reference's dist_util / logger / video_datasets /
scripts read from MPI.COMM_WORLD.  Test infrastructure (oracle/), used only to import the reference's own files."""


class _Comm:
    rank, size = 0, 1

    def Get_rank(self):
        return 0

    def Get_size(self):
        return 1

    def bcast(self, x, root=0):
        return x

    def gather(self, x, root=0):
        return [x]

    def allgather(self, x):
        return [x]

    def Barrier(self):
        pass


class _MPI:
    COMM_WORLD = _Comm()


MPI = _MPI()
