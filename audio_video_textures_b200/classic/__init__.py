"""Drop-in replacements for baselines/classic_video_textures/{computeD1,computeD2,q_learning,video_textures}.py.

Same module names, function names, argument order/defaults and returned tuples of CUDA fp32 tensors
as the reference; the arithmetic runs in libavtex.so (hand-written sm_100a CUDA).  Put this directory
on sys.path (or `from audio_video_textures_b200.classic import computeD1`) instead of the reference's.
"""
