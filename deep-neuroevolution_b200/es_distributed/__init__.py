"""Reference-facing API: the same module and function names as uber-research/deep-neuroevolution's
``es_distributed`` package (es, ga, nses, policies, optimizers, main), backed by libdne.so.

``configurations/*.json`` of the reference drive it unchanged:  python -m es_distributed.main master --algo es --exp_file ...
"""
