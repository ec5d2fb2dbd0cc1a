"""`miRge2.0 annotate` command line (src/mirge/utils/parseArgument.py:31-53,
src/mirge/__main__.py:73-392), served by the CUDA pipeline.

Same flags, same library-directory layout, same output directory
`<-o>/miRge.<YYYY-MM-DD_HH-MM-SS>/` and table files.  `-pb` (bowtie directory)
and `-cpu` are accepted and ignored: no bowtie or cutadapt process is started.
Only the annotate sub-command exists; -gff writes <sample>_isomiRs.gff (mirge_b200/gff.py); -ai writes the A-to-I editing report (mirge_b200/a2i.py); -trf outputs
are not produced (DESIGN.md "Out of scope").
"""
import argparse
import os
import sys
import time

VERSION = '2.0'


def parse_arguments(argv=None):
    parser = argparse.ArgumentParser(prog='miRge2.0', formatter_class=argparse.RawTextHelpFormatter,
                                     description='miRge2.0 (annotate mode) on NVIDIA B200')
    sub = parser.add_subparsers(dest='command')
    p = sub.add_parser('annotate')
    p.add_argument('-s', nargs='*', required=True, dest='sampleList', metavar='sample <required>')
    p.add_argument('-o', default=os.getcwd(), dest='output_dir', metavar='<dir>')
    p.add_argument('-d', default='miRBase', dest='miRNA_database', metavar='<string required>')
    p.add_argument('-pb', required=False, default='', dest='bowtieBinary', metavar='<dir>', help='accepted and ignored')
    p.add_argument('-lib', required=True, dest='libraryPath', metavar='<dir required>')
    p.add_argument('-sp', required=True, dest='species', metavar='<string required>')
    p.add_argument('-ex', default='0.1', dest='canoRatio', metavar='<float>')
    p.add_argument('-ad', default='none', dest='adapter', metavar='<string>')
    p.add_argument('-phred64', action='store_true')
    p.add_argument('-spikeIn', dest='spikeIn', action='store_true')
    p.add_argument('-tcf', dest='trimmed_collapsed_fa', action='store_true')
    p.add_argument('-di', dest='diff_isomirs', action='store_true')
    p.add_argument('-cpu', dest='cpu', metavar='<int>', default='1', help='accepted and ignored')
    p.add_argument('-ai', dest='a_to_i', action='store_true')
    p.add_argument('-gff', dest='gff_output', action='store_true')
    p.add_argument('-trf', dest='trf_output', action='store_true')
    p.add_argument('--gpu', dest='gpu', type=int, default=0, help='CUDA device index')
    p.add_argument('--max-unique', dest='max_unique', type=int, default=None, help='capacity of the unique-sequence table')
    p.add_argument('--version', action='version', version=VERSION)
    return parser.parse_args(argv)


def resolve_samples(sample_args):
    """-s: a list of *.fastq / *.fastq.gz files, or one text file of paths (__main__.py:289-314)."""
    def is_fastq(p):
        b = os.path.basename(p).split('.')
        return b[-1] == 'fastq' or '.'.join(b[-2:]) == 'fastq.gz'
    if all(is_fastq(p) for p in sample_args):
        for p in sample_args:
            if not os.path.isfile(os.path.abspath(p)):
                print("%s can't be found in current directory. Please check it." % p)
                sys.exit(1)
        return list(sample_args)
    if len(sample_args) == 1:
        out = []
        try:
            with open(sample_args[0]) as f:
                for line in f:
                    s = line.strip()
                    if s not in out:
                        if os.path.isfile(os.path.abspath(s)):
                            out.append(s)
                        else:
                            print('%s cannot be found, please check the path of the sample file.' % s)
                            sys.exit(1)
        except IOError:
            print('%s is not a file, please check it.' % sample_args[0])
            sys.exit(1)
        return out
    print("The format of input argument '-s' is wrong, please check it.")
    sys.exit(1)


def main(argv=None):
    args = parse_arguments(argv)
    if args.command != 'annotate':
        print('Only `miRge2.0 annotate` is implemented by this build.', file=sys.stderr)
        sys.exit(1)
    db = args.miRNA_database.lower()
    if db not in ('mirbase', 'mirgenedb'):
        print("The value of parameter '-d' is invalid. Please check it", file=sys.stderr)
        sys.exit(1)
    db = 'miRBase' if db == 'mirbase' else 'MirGeneDB'
    adapter = args.adapter
    if adapter == 'illumina':
        adapter = 'TGGAATTCTCGGGTGCCAAGGAACTCCAG'      # __main__.py:124-125
    elif adapter == 'ion':
        adapter = '11'                                  # __main__.py:126-127 (cutadapt rejects it: starts with a digit)
    for flag, name in ((args.trf_output, '-trf'),):
        if flag:
            print('warning: %s output is not produced by this build' % name, file=sys.stderr)

    from .libs import LibraryError
    from .gff import GffError
    from .pipeline import AnnotateRun
    from .tables import ZeroMiRNAError
    from .engine import MirgeError

    samples = resolve_samples(args.sampleList)
    # launched under torch.distributed.run: one process per GPU, the file chunks of every sample shard over the ranks
    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    device = args.gpu
    log = print
    stamp = time.strftime('%Y-%m-%d_%H-%M-%S', time.localtime())
    if world > 1:
        import torch
        import torch.distributed as dist
        device = int(os.environ.get('LOCAL_RANK', rank))
        torch.cuda.set_device(device)
        dist.init_process_group('nccl', device_id=torch.device('cuda', device))
        box = [stamp]
        dist.broadcast_object_list(box, src=0)          # one output directory name for the job
        stamp = box[0]
        if rank != 0:
            log = lambda *a, **k: None                  # noqa: E731
    outdir = os.path.join(os.path.abspath(args.output_dir), 'miRge.' + stamp)
    t0 = time.time()
    try:
        run = AnnotateRun(samples, args.libraryPath, args.species, db, adapter, args.spikeIn,
                          max_unique=args.max_unique, device=device, log=log, rank=rank, world=world)
        if args.a_to_i and not run.libdir.has_genome():
            raise LibraryError('The bowtie index file of %s_genome.*.ebwt is not located at %s, please check it (-ai needs the genome).'
                               % (args.species, run.libdir.index_path))
        if args.gff_output and not os.path.isfile(run.libdir.gff3):     # extractPreMiRName opens it unguarded (IOError upstream)
            raise LibraryError('%s does not exist: -gff needs the miRNA coordinate file.' % run.libdir.gff3)
        if rank == 0:
            os.mkdir(outdir)
        run.run()
        t1 = time.time()
        log('Summarizing and tabulating results...')
        run.write_outputs(outdir, args.canoRatio, tcf=args.trimmed_collapsed_fa, isomir_diff=args.diff_isomirs, gff=args.gff_output, a_to_i=args.a_to_i)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
            if rank != 0:
                return outdir
    except (LibraryError, GffError) as e:
        print(str(e))
        sys.exit(1)
    except ZeroMiRNAError as e:
        print(str(e))
        sys.exit(1)
    except MirgeError as e:
        print(str(e), file=sys.stderr)
        sys.exit(1)
    t2 = time.time()
    print('Summary Complete (%.2f sec)' % (t2 - t1))
    print('Annotation of miRge2.0 Completed (%.2f sec)' % (t2 - t0))
    return outdir


if __name__ == '__main__':
    main()
