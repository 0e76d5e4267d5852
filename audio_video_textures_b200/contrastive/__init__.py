"""Drop-in pieces for contrastive_video_textures/{models/models.py, validate.py, main.py}: the
similarity tail, the selection block and the `-e` synthesis loop at the embedding boundary."""
