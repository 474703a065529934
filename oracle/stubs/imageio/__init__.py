"""Empty stand-in for imageio (absent from the image; only the reference's mp4/gif writers call into it)."""
