"""Constants of realtime/config.py:15-56 that shape the hot path."""
SR = 96000
BLOCKSIZE = 128
N_CHANNELS = 3
N_FFT = 2048
HOP_LENGTH = BLOCKSIZE
