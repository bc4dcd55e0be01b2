// Edge.h -- drop-in for MetaGenomics/Edge.h:16-61: one directed overlap edge with its twin pointer
// and the (initially empty) composite lists the downstream contraction fills.
#ifndef OGB_HOST_EDGE_H_
#define OGB_HOST_EDGE_H_

#include "Common.h"
#include "Read.h"

class Edge
{
	private:
		Read *source;
		Read *destination;
		UINT8 overlapOrientation;				// 0 u<--<v  1 u<-->v  2 u>--<v  3 u>-->v
		UINT64 overlapOffset;					// start of v relative to u
		vector<UINT64> *listOfReads;
		vector<UINT16> *listOfOverlapOffsets;
		vector<UINT8> *listOfOrientations;
		Edge *reverseEdge;

	public:
		bool transitiveRemovalFlag;
		UINT16 flow;
		UINT64 coverageDepth;
		UINT64 SD;
		Edge(void);
		Edge(Read *from, Read *to, UINT64 orient, UINT64 length);
		Edge(Read *from, Read *to, UINT64 orient, UINT64 length, vector<UINT64> *listReads, vector<UINT16> *listOverlapOffsets, vector<UINT8> *listOrientations);
		~Edge();
		bool makeEdge(Read *from, Read *to, UINT64 orient, UINT64 length);
		bool makeEdge(Read *from, Read *to, UINT64 orient, UINT64 length, vector<UINT64> *listReads, vector<UINT16> *listOverlapOffsets, vector<UINT8> *listOrientations);
		bool setReverseEdge(Edge *edge);
		Read *getSourceRead() { return source; }
		Read *getDestinationRead() { return destination; }
		UINT8 getOrientation() { return overlapOrientation; }
		UINT64 getOverlapOffset() { return overlapOffset; }
		vector<UINT64> *getListOfReads() { return listOfReads; }
		vector<UINT16> *getListOfOverlapOffsets() { return listOfOverlapOffsets; }
		vector<UINT8> *getListOfOrientations() { return listOfOrientations; }
		Edge *getReverseEdge() { return reverseEdge; }
};

#endif
