#!/usr/bin/env python
"""Randomised model check of the hand-over protocol of the experimental warp-specialised slab kernel
(jots_b200/csrc/apply_slab5.cuh): two sequential agents (pencil group, plane group), four gather-row slots, one staging
buffer, two plane buffers, the full / done mbarriers as phase counters, asynchronous copies that complete at random times
between issue and wait.  Asserts that nothing is read before it is ready or overwritten before its last reader, and that
no interleaving deadlocks, for 0..11 tiles per block.  CPU only: python scripts/slab5_protocol_model.py"""
import random
def run(n, seed):
    rnd=random.Random(seed)
    # resources
    idx=[None]*4      # which tile's rows are in slot (or ('inflight',tile))
    stage=None        # tile whose values are staged, or ('inflight', tile)
    planes=[None,None]  # state: ('p1',tile) | ('res',tile) | None
    full=[0,0]; done=[0,0]   # completed phase counts
    pending=[]        # async copies: (kind, tile)
    log=[]
    def complete_some(all_=False):
        nonlocal stage
        keep=[]
        for kind,tile in pending:
            if all_ or rnd.random()<0.5:
                if kind=='idx':
                    assert idx[tile%4]==('inflight',tile); idx[tile%4]=tile
                else:
                    assert stage==('inflight',tile); stage=tile
            else: keep.append((kind,tile))
        pending[:]=keep
    def issue_idx(i):
        if i<n:
            # slot must not hold rows still needed: tiles > i-4 that are not finished with phase 3
            cur=idx[i%4]
            assert cur is None or (isinstance(cur,int) and cur in p3_done), ('idx overwrite',i,cur)
            idx[i%4]=('inflight',i); pending.append(('idx',i))
    def issue_vals(i):
        nonlocal stage
        if i<n:
            assert idx[i%4]==i, ('vals needs rows',i,idx)
            assert stage is None or (isinstance(stage,int) and stage in p1_done), ('stage overwrite',i,stage)
            stage=('inflight',i); pending.append(('vals',i))
    p1_done=set(); p3_done=set(); p2_done=set()
    # generators for the two groups
    def pencil():
        nonlocal stage
        for i in range(4): issue_idx(i)
        yield 'wait0'
        issue_vals(0); yield 'wait0'
        if n>0:
            assert stage==0 and idx[0]==0
            assert planes[0] is None
            planes[0]=('p1',0); p1_done.add(0); full[0]+=1
        yield 'bar'
        issue_vals(1); yield 'wait0'
        for it in range(n):
            if it+1<n:
                assert stage==it+1 and idx[(it+1)%4]==it+1, ('p1 inputs',it,stage,idx)
                b=(it+1)%2
                assert planes[b] is None or planes[b]==('consumed',it-1), ('p1 overwrite',it,planes)
                planes[b]=('p1',it+1); p1_done.add(it+1); full[b]+=1
            yield 'bar'
            issue_vals(it+2)
            # wait done[it%2] phase (it//2): completed count must reach it//2+1
            while done[it%2] < it//2+1: yield 'spin'
            assert planes[it%2]==('res',it), ('p3 input',it,planes)
            assert idx[it%4]==it
            planes[it%2]=('consumed',it); p3_done.add(it)
            yield 'wait0'
            issue_idx(it+4)
        yield 'wait0'
    def plane():
        for it in range(n):
            while full[it%2] < it//2+1: yield 'spin'
            assert planes[it%2]==('p1',it), ('p2 input',it,planes)
            yield 'work'
            planes[it%2]=('res',it); p2_done.add(it); done[it%2]+=1
    gens=[pencil(),plane()]; alive=[True,True]; steps=0
    while any(alive):
        steps+=1
        assert steps<100000,'deadlock?'
        g=rnd.randrange(2)
        if not alive[g]: continue
        try:
            ev=next(gens[g])
            if ev=='wait0': complete_some(all_=True)
            else: complete_some()
        except StopIteration:
            alive[g]=False
    assert p3_done==set(range(n)) and p2_done==set(range(n))
for n in range(0,12):
    for seed in range(200): run(n,seed)
print('protocol ok')
