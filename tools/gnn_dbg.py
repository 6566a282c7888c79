import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from conftest import random_ksat
from oracle import gnn_oracle as go
import sls_sat_solving_with_deep_learning_b200 as eng
from sls_sat_solving_with_deep_learning_b200 import gnn
import __graft_entry__ as g; g.build()
eng.set_device(0)
n,m=60,250
cl=random_ksat(n,m,3,seed=5); graph=go.lcg_graph(n,cl)
for layers,steps in (((32,32),1),((200,200),1),((200,200),5)):
    params=go.init_params(seed=42,mlp_layers=layers,num_steps=steps,bias_scale=0.2)
    ref=go.forward(params,graph,dtype=np.float64,num_steps=steps)
    net=gnn.InteractionNetworkLCG(params,num_message_passing_steps=steps)
    dec=net.apply(graph)
    print(layers,steps,'max err decoded',np.abs(dec-ref).max(),'ref range',ref.min(),ref.max(),'ms',net.last_kernel_ms, 'nan',np.isnan(dec).sum())
    print(' p err', np.abs(gnn.get_model_probabilities(dec,n)-go.model_probabilities(ref,n)).max())
    print(dec[:4].ravel(), ref[:4].ravel())
