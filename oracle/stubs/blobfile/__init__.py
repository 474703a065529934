"""Local-filesystem stand-in for blobfile (absent from the image): the four calls the reference makes (train_util.py,
dist_util.py).  Test infrastructure (oracle/)."""
import os

BlobFile = open
join = os.path.join
dirname = os.path.dirname
exists = os.path.exists


def listdir(p):
    return os.listdir(p)
