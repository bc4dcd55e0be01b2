// Read.h -- host view of one unique read (drop-in for MetaGenomics/Read.h:31-72).
//
// The bases live once, 2-bit packed, in the Dataset's store (and in HBM); a Read keeps only its
// identity, the hot-path output superReadID (Read.h:50) and the bookkeeping lists the downstream
// host stages append to. getStringForward/Reverse decode on demand.
#ifndef OGB_HOST_READ_H_
#define OGB_HOST_READ_H_

#include "Common.h"

class Edge;
class Dataset;

struct MPlist
{
	UINT64 matePairID;				// ID of the mate
	UINT8 matePairOrientation;		// bit1 = this read forward, bit0 = mate forward (Read.h:18-22)
	UINT8 datasetNumber;
};

class Read
{
	private:
		UINT64 readNumber;
		UINT32 frequency;
		UINT16 length;
		const Dataset *owner;			// packed bases are fetched from here
		vector<MPlist> *matePairList;
		vector<Edge *> *listOfEdgesForward;
		vector<UINT64> *locationOnEdgeForward;
		vector<Edge *> *listOfEdgesReverse;
		vector<UINT64> *locationOnEdgeReverse;
		friend class Dataset;

	public:
		UINT64 coverageDepth;
		UINT64 locationInDataset;
		bool isContainedRead;			// never written by the reference either
		UINT64 superReadID;				// 0 = not contained, else ID of the longest containing read

		Read(void);
		~Read(void);

		bool setReadNumber(UINT64 id);
		bool setFrequency(UINT32 freq);
		string getStringForward(void) const;
		string getStringReverse(void) const;
		UINT16 getReadLength(void) const { return length; }
		UINT64 getReadNumber(void) const { return readNumber; }
		UINT32 getFrequency(void) const { return frequency; }
		vector<MPlist> *getMatePairList(void) { return matePairList; }
		vector<Edge *> *getListOfEdgesForward(void) { return listOfEdgesForward; }
		vector<UINT64> *getLocationOnEdgeForward(void) { return locationOnEdgeForward; }
		vector<Edge *> *getListOfEdgesReverse(void) { return listOfEdgesReverse; }
		vector<UINT64> *getLocationOnEdgeReverse(void) { return locationOnEdgeReverse; }
		bool addMatePair(Read *r, UINT8 orientation, UINT64 datasetNumber);
};

#endif
