"""waveform_ot_b200 -- B200 (sm_100a) implementation of waveform-ot's
fingerprint + marginal-Wasserstein misfit hot path behind the reference's own
Python surface (FingerprintLib.waveformFP, OTlib.OTpdf / wasser / MargWasserstein).

Importing the compute modules requires the in-tree CUDA library
(waveform_ot_b200/libwfot.so, built by `python -m waveform_ot_b200.build`);
there is no CPU fallback.
"""
__version__ = "0.1.0"
