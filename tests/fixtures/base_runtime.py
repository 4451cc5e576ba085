# a tiny `_base_` file in the style of the reference's configs/_base_/default_runtime.py
log_config = dict(interval=50, hooks=[dict(type='TextLoggerHook', by_epoch=False)])
dist_params = dict(backend='nccl')
optimizer = dict(type='SGD', lr=0.01, momentum=0.9, weight_decay=0.0005)
