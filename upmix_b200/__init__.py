"""upmix_b200: B200-native multi-band STFT centre extraction (drop-in for willleskowitz/upmix's
python-prototype/center_extraction.py + main.py hot path)."""
__version__ = "0.1.0"
